"""ctypes binding of libmmfn_b200.so.  Prototypes are read from include/mmfn_b200.h so the
Python side can never drift from the C ABI.  There is no fallback: a missing library or a
failing call raises."""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "mmfn_b200.h")
LIB_PATH = os.path.join(HERE, "lib", "libmmfn_b200.so")

_CTYPES = {
    "int": ctypes.c_int, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64,
    "float": ctypes.c_float, "double": ctypes.c_double, "cudaStream_t": ctypes.c_void_p,
}
_PROTO = re.compile(r"^int\s+(mmfn_\w+)\s*\(([^)]*)\)\s*;", re.M)


def parse_header(path=HEADER):
    """-> {name: [(ctype, argname), ...]} for every function the header declares."""
    protos = {}
    for name, args in _PROTO.findall(open(path).read()):
        sig = []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    sig.append((ctypes.c_void_p, a.split("*")[-1].strip()))
                else:
                    toks = [t for t in a.split() if t != "const"]
                    sig.append((_CTYPES[toks[0]], toks[1]))
        protos[name] = sig
    return protos


class MmfnError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise MmfnError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU or eager fallback)")
        self._dll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        self.launches = 0
        for name, sig in self.protos.items():
            fn = getattr(self._dll, name)      # AttributeError if the .so lacks a declared symbol
            fn.restype = ctypes.c_int
            fn.argtypes = [t for t, _ in sig]
            if name not in ("mmfn_version", "mmfn_last_error"):
                setattr(self, name[len("mmfn_"):], self._wrap(name, fn))
        self.version = self._dll.mmfn_version()

    # ---- process-wide dropout offset -------------------------------------------------------------
    # mmfn_rng_bind stores a DEVICE ADDRESS in constant memory: it must outlive every engine and must not be rebound
    # by a second engine, so the library owns ONE int64 cell per device that is never freed.
    _rng = {}

    def rng_tensor(self, device, rank=0):
        import torch
        device = torch.device(device)
        key = device.index if device.index is not None else torch.cuda.current_device()
        t = self._rng.get(key)
        if t is None:
            # ranks start from different offsets: data-parallel replicas must not draw identical masks
            t = self._rng[key] = torch.full((1,), int(rank) * 7919 * 1000003, device=device, dtype=torch.int64)
            with torch.cuda.device(device):
                self.rng_bind(t.data_ptr())
        return t

    def last_error(self):
        buf = ctypes.create_string_buffer(512)
        self._dll.mmfn_last_error(buf, 512)
        return buf.value.decode()

    # kernels launched per C-ABI call when it is not exactly one (for the bench's launch count)
    KERNELS_PER_CALL = {"mmfn_bn_train_fwd": 2, "mmfn_bn_eval_fwd": 2, "mmfn_bn_train_bwd": 2,
                        "mmfn_adamw_step": 2, "mmfn_tokens_bwd": 2, "mmfn_bev_scatter_ws": 2,
                        "mmfn_stem_bn_relu_maxpool_fwd": 2, "mmfn_stem_bn_relu_maxpool_bwd": 2}

    def _wrap(self, name, fn):
        nk = self.KERNELS_PER_CALL.get(name, 1)
        short = name[len("mmfn_"):]

        def call(*args):
            prof = self.profile
            if prof is not None:
                import torch
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = fn(*args)
                e1.record()
                prof.append((short, self.next_work, e0, e1))
            else:
                rc = fn(*args)
            self.next_work = None
            self.launches += nk
            if rc != 0:
                raise MmfnError(f"{name} failed (code {rc}): {self.last_error()}")
        call.__name__ = name
        return call

    # ---- optional per-call device timing (bench.py roofline leg) -------------------------------
    profile = None        # list of (fn, (flops, bytes) | None, start_event, end_event) while enabled
    next_work = None      # set by ops wrappers right before a call: algorithmic (flops, bytes)

    @staticmethod
    def kernel_class(fn, work):
        """conv | linear | attention | fused_gpt (tensor-pipe bound) | hbm (everything else: normalisation, pooling, scatter,
        optimizer ...).  Batched GEMMs are the attention products (QK^T, PV and their gradients)."""
        if fn.startswith("conv2d_"):
            return "conv"
        if fn.startswith("attention_"):
            return "attention"
        if fn.startswith("gpt_small_") and fn != "gpt_small_transpose":
            return "fused_gpt"          # whole-GPT forward / row-local backward (linears + attention of the narrow transformers)
        if fn.startswith("gemm_"):
            return "attention" if (work and len(work) >= 6 and work[5] > 1) else "linear"
        return "hbm"

    def start_profile(self):
        self.profile = []

    def stop_profile(self):
        """-> {fn: dict(calls, ms, flops, bytes)}; caller must have synchronised the device."""
        out, shapes, classes = {}, {}, {}
        for fn, work, e0, e1 in self.profile or []:
            ms = e0.elapsed_time(e1)
            c = classes.setdefault(self.kernel_class(fn, work), dict(calls=0, ms=0.0, flops=0.0))
            c["calls"] += 1
            c["ms"] += ms
            c["flops"] += work[0] if work else 0.0
            d = out.setdefault(fn, dict(calls=0, ms=0.0, flops=0.0, bytes=0.0))
            d["calls"] += 1
            d["ms"] += ms
            if work:
                d["flops"] += work[0]
                d["bytes"] += work[1]
                s = shapes.setdefault((fn,) + tuple(work[2:] if len(work) > 2 else work[:1]), [0, 0.0, work[0]])
                s[0] += 1
                s[1] += ms
        self.profile = None
        self.last_classes = classes
        self.last_shapes = sorted(([k, v[0], v[1], v[2]] for k, v in shapes.items()), key=lambda r: -r[2])
        return out


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
