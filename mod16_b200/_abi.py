'''
ctypes binding of ``include/mod16_b200.h`` -- the only way the Python layer reaches the
CUDA kernels.  There is no CPU fallback: if ``libmod16_b200.so`` is missing or CUDA is
unavailable, the first call raises.
'''
import ctypes
import os

N_PARAMS = 11
N_PFT = 13
N_DRIVERS = 14

# flag words (mirror include/mod16_b200.h)
SEM_STATIC, SEM_INSTANCE = 0x0, 0x1
OUT_SUM, OUT_DAY_NIGHT, OUT_COMPONENTS, OUT_ANNUAL, OUT_POTENTIAL = 0x00, 0x10, 0x20, 0x30, 0x40
OUT_COMPOSITE, OUT_SUM_I16 = 0x50, 0x60
GATE_GLOBAL, GATE_ELEMENT, GATE_OPEN, GATE_CLOSED = 0x000, 0x100, 0x200, 0x300
GATE_OPEN_TRACK, GATE_FIXUP = 0x400, 0x500
FORCE_GENERIC, FORCE_MIRROR = 0x1000, 0x2000
KERNEL_PIPE, KERNEL_STRIDED, KERNEL_EXACT, KERNEL_FAST64 = 1, 2, 3, 4  # last_forward_kernel & 0xff
T_F32, T_I16, T_U8 = 0, 1, 2                            # mod16_raw_driver_t.dtype
PREP_VPD_NIGHT_CLAMP, PREP_ELEVATION, PREP_VPD_FROM_QV = 0x1, 0x2, 0x4
ERR_LAYOUT = -7
ANNUAL_OVERFLOW = 0x40000000
GATE_FLAGS = {'global': GATE_GLOBAL, 'element': GATE_ELEMENT, 'open': GATE_OPEN,
              'closed': GATE_CLOSED, 'open_track': GATE_OPEN_TRACK, 'fixup': GATE_FIXUP}

ABI_VERSION = 12
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib',
                        'libmod16_b200.so')


class Driver(ctypes.Structure):
    _fields_ = [('ptr', ctypes.c_void_p), ('t_stride', ctypes.c_int64),
                ('n_stride', ctypes.c_int64), ('scalar', ctypes.c_double)]


class Drivers(ctypes.Structure):
    _fields_ = [('drv', Driver * N_DRIVERS), ('r_corr', Driver * 2),
                ('has_r_corr', ctypes.c_int32), ('f32_mask', ctypes.c_uint32)]


class RawDriver(ctypes.Structure):
    _fields_ = [('ptr', ctypes.c_void_p), ('t_stride', ctypes.c_int64),
                ('n_stride', ctypes.c_int64), ('scalar', ctypes.c_double),
                ('dtype', ctypes.c_int32), ('has_fill', ctypes.c_int32),
                ('scale', ctypes.c_double), ('offset', ctypes.c_double),
                ('fill', ctypes.c_int64)]


class RawDrivers(ctypes.Structure):
    _fields_ = [('drv', RawDriver * N_DRIVERS), ('prep', ctypes.c_uint32),
                ('reserved', ctypes.c_uint32), ('ps', RawDriver * 2)]


class Params(ctypes.Structure):
    _fields_ = [('pft', ctypes.c_void_p),
                ('table', (ctypes.c_double * N_PFT) * N_PARAMS),
                ('arrays', ctypes.c_void_p * N_PARAMS),
                ('array_stride', ctypes.c_int64 * N_PARAMS)]


class Out(ctypes.Structure):
    _fields_ = [('out', ctypes.c_void_p * 6), ('gate_flag', ctypes.c_void_p),
                ('workspace', ctypes.c_void_p), ('workspace_bytes', ctypes.c_int64),
                ('group', ctypes.c_void_p), ('n_groups', ctypes.c_int32),
                ('reserved', ctypes.c_int32), ('day_weight', ctypes.c_void_p),
                ('night_weight', ctypes.c_void_p), ('out_scale', ctypes.c_double)]


class Aggregate(ctypes.Structure):
    _fields_ = [('x', ctypes.c_void_p), ('divisor', ctypes.c_void_p), ('scale', ctypes.c_double),
                ('group', ctypes.c_void_p), ('n_groups', ctypes.c_int32),
                ('reserved', ctypes.c_int32), ('out', ctypes.c_void_p),
                ('total', ctypes.c_void_p)]


AGG_CLAMP_NEG, AGG_SKIP_NAN = 0x1, 0x2


class Objective(ctypes.Structure):
    _fields_ = [('draws', ctypes.c_void_p), ('observed', Driver), ('weights', Driver),
                ('sse', ctypes.c_void_p), ('count', ctypes.c_void_p),
                ('gate', ctypes.c_void_p), ('workspace', ctypes.c_void_p),
                ('workspace_bytes', ctypes.c_int64), ('raw', ctypes.c_void_p)]


MAX_OP_INPUTS = 12
OPS = ('svp', 'svp_slope', 'lhv', 'psychrometric', 'air_density', 'rhumidity', 'vpd',
       'air_pressure', 'radiation_net', 'soil_heat_flux', 'radiation_soil',
       'potential_soil_evaporation', 'evaporation_soil', 'evaporation_wet_canopy',
       'surface_conductance', 'transpiration', 'potential_transpiration',
       'daylength_composite')


class Elementwise(ctypes.Structure):
    _fields_ = [('op', ctypes.c_int32), ('present', ctypes.c_uint32),
                ('inp', Driver * MAX_OP_INPUTS), ('params', ctypes.c_double * N_PARAMS),
                ('opt', ctypes.c_double * 2), ('out', ctypes.c_void_p * 2)]


class Mod16Error(RuntimeError):
    pass


class LayoutError(Mod16Error):
    'MOD16_ERR_LAYOUT: a storage-format layout the fused kernel cannot stream.'


_lib = None

# every symbol include/mod16_b200.h declares; tests/test_abi_symbols.py checks the list
# against the header and the built library
EXPORTS = {
    'mod16_abi_version': (ctypes.c_int, []),
    'mod16_workspace_bytes': (ctypes.c_int64, []),
    'mod16_last_cuda_error': (ctypes.c_int, []),
    'mod16_last_forward_kernel': (ctypes.c_int, []),
    'mod16_error_string': (ctypes.c_char_p, [ctypes.c_int]),
    'mod16_forward_raw_f32': (ctypes.c_int, [
        ctypes.POINTER(RawDrivers), ctypes.POINTER(Params), ctypes.POINTER(Out),
        ctypes.c_int64, ctypes.c_int64, ctypes.c_uint32, ctypes.c_void_p]),
    'mod16_forward_f32': (ctypes.c_int, [
        ctypes.POINTER(Drivers), ctypes.POINTER(Params), ctypes.POINTER(Out),
        ctypes.c_int64, ctypes.c_int64, ctypes.c_uint32, ctypes.c_void_p]),
    'mod16_forward_f64': (ctypes.c_int, [
        ctypes.POINTER(Drivers), ctypes.POINTER(Params), ctypes.POINTER(Out),
        ctypes.c_int64, ctypes.c_int64, ctypes.c_uint32, ctypes.c_void_p]),
    'mod16_objective_workspace_bytes': (ctypes.c_int64, [
        ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]),
    'mod16_objective_batch_f32': (ctypes.c_int, [
        ctypes.POINTER(Drivers), ctypes.POINTER(Objective), ctypes.c_int64, ctypes.c_int64,
        ctypes.c_int64, ctypes.c_uint32, ctypes.c_void_p]),
    'mod16_objective_batch_f64': (ctypes.c_int, [
        ctypes.POINTER(Drivers), ctypes.POINTER(Objective), ctypes.c_int64, ctypes.c_int64,
        ctypes.c_int64, ctypes.c_uint32, ctypes.c_void_p]),
    'mod16_elementwise_f32': (ctypes.c_int, [
        ctypes.POINTER(Elementwise), ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]),
    'mod16_elementwise_f64': (ctypes.c_int, [
        ctypes.POINTER(Elementwise), ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]),
    'mod16_time_aggregate_f32': (ctypes.c_int, [
        ctypes.POINTER(Aggregate), ctypes.c_int64, ctypes.c_int64, ctypes.c_uint32,
        ctypes.c_void_p]),
    'mod16_time_aggregate_f64': (ctypes.c_int, [
        ctypes.POINTER(Aggregate), ctypes.c_int64, ctypes.c_int64, ctypes.c_uint32,
        ctypes.c_void_p]),
}


def load(path=None):
    'Loads the shared library (once) and declares the prototypes.'
    global _lib
    if _lib is not None and path is None:
        return _lib
    # MOD16_B200_LIB: an alternative build of the same ABI (kernel A/B experiments)
    path = path or os.environ.get('MOD16_B200_LIB') or LIB_PATH
    if not os.path.exists(path):
        raise Mod16Error(
            'CUDA library %s is not built; run `python -m mod16_b200.build` '
            '(there is no CPU fallback)' % path)
    if path == LIB_PATH:
        from . import build as _build
        if _build.is_stale():
            raise Mod16Error(
                '%s was not built from the sources in this tree; run `python -m '
                'mod16_b200.build`' % path)
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in EXPORTS.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.mod16_abi_version() != ABI_VERSION:
        raise Mod16Error('ABI version mismatch: library %d, binding %d' % (
            lib.mod16_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        lib = load()
        msg = lib.mod16_error_string(rc).decode()
        if rc == -4:
            msg += ' [cudaError %d]' % lib.mod16_last_cuda_error()
        raise (LayoutError if rc == ERR_LAYOUT else Mod16Error)(
            'mod16_b200: %s (code %d)' % (msg, rc))
