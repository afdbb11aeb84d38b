"""ctypes binding of libngan_b200.so.  The prototypes are parsed from include/ngan_b200.h so the binding can
never drift from the declared C ABI.  Loading fails loudly when the library has not been built."""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('NGAN_LIB') or os.path.join(_HERE, 'libngan_b200.so')   # NGAN_LIB: debug builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), 'include', 'ngan_b200.h')

_CTYPES = {'int': ctypes.c_int, 'float': ctypes.c_float, 'long long': ctypes.c_longlong}


class AdamTensor(ctypes.Structure):
    """Mirror of ngan_adam_tensor."""
    _fields_ = [('p', ctypes.c_void_p), ('g', ctypes.c_void_p), ('m', ctypes.c_void_p), ('v', ctypes.c_void_p),
                ('shadow_bf16', ctypes.c_void_p), ('n', ctypes.c_longlong), ('step_size', ctypes.c_float),
                ('inv_bc2_sqrt', ctypes.c_float), ('shadow_k', ctypes.c_int), ('shadow_c', ctypes.c_int),
                ('shadow_ss', ctypes.c_int), ('shadow_kind', ctypes.c_int), ('dyn', ctypes.c_void_p)]


def parse_header(path=HEADER_PATH):
    """Returns {name: (restype, [(ctype, argname), ...])} for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'\b(int|long long|const char\*)\s+(ngan_\w+)\s*\(([^)]*)\)\s*;', src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        parsed = []
        if args and args != 'void':
            for a in args.split(','):
                a = ' '.join(a.split())
                if '*' in a:
                    parsed.append((ctypes.c_void_p, a.split('*')[-1].strip()))
                else:
                    ty, nm = a.rsplit(' ', 1)
                    parsed.append((_CTYPES[ty], nm))
        protos[name] = ({'int': ctypes.c_int, 'long long': ctypes.c_longlong}.get(ret, ctypes.c_char_p), parsed)
    return protos


class NganError(RuntimeError):
    pass


_lib = None
_protos = None


def load():
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                          '(neuron_gan_b200 has no CPU or PyTorch fallback)')
    lib = ctypes.CDLL(LIB_PATH)
    _protos = parse_header()
    for name, (restype, args) in _protos.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = [t for t, _ in args]
    _lib = lib
    return lib


_QUERIES = {'ngan_linear_fwd_workspace_bytes', 'ngan_augment_workspace_bytes', 'ngan_conv_weight_is_folded', 'ngan_version',
            'ngan_conv3x3_wgrad_workspace_bytes', 'ngan_pixel_reduction_workspace_bytes',
            'ngan_gp_loss_workspace_bytes', 'ngan_similarity_loss_workspace_bytes'}   # no launch, value returned
# kernels launched per C-ABI call (everything not listed launches exactly one; callers pass `launches=` where the
# count depends on the arguments, e.g. one reduction per requested parameter gradient)
_LAUNCHES = {'ngan_gp_loss': 2, 'ngan_similarity_loss': 2, 'ngan_augment_batch': 3, 'ngan_conv3x3_wgrad': 2, 'ngan_bias_grad': 2,
             'ngan_memset': 0}
launch_count = 0          # running count of kernels launched through this binding (bench.py reads it)
_profile = None           # when a list: (name, int args, start event, end event) per call


def start_profile():
    """Record a CUDA event pair around every call (bench.py's per-kernel timing pass; perturbs throughput)."""
    global _profile
    _profile = []


def stop_profile():
    """Returns [(name, int-args tuple, milliseconds)] and disables profiling. Synchronises."""
    global _profile
    import torch
    torch.cuda.synchronize()
    out = [(n, a, e0.elapsed_time(e1)) for n, a, e0, e1 in _profile]
    _profile = None
    return out


def call(name, *args, launches=None):
    global launch_count
    lib = load()
    if _profile is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(lib, name)(*args)
    if name in _QUERIES:
        return rc
    if rc != 0:
        raise NganError(f'{name} failed ({rc}): {lib.ngan_last_error().decode()}')
    launch_count += _LAUNCHES.get(name, 1) if launches is None else launches
    if _profile is not None:
        e1.record()
        _profile.append((name, tuple(a for a in args if isinstance(a, int)), e0, e1))
