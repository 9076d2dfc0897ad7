"""ctypes binding of libvcagan_b200.so.  Prototypes are parsed from include/vcagan.h so the header is the
single source of truth for the C ABI.  There is no fallback: if the library is missing or a call fails this
raises, it never silently routes to torch or the CPU."""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
LIB_PATH = os.path.join(_HERE, "libvcagan_b200.so")
HEADER = os.path.join(_ROOT, "include", "vcagan.h")


class ConvGeom(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in
                ("N", "ID", "IH", "IW", "Cin", "OD", "OH", "OW", "Cout", "KD", "KH", "KW", "sd", "sh", "sw", "pd", "ph", "pw")]

    def key(self):
        return tuple(getattr(self, f) for f, _ in self._fields_)


_CT = {
    "int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double, "long long": ctypes.c_longlong,
    "unsigned long long": ctypes.c_ulonglong, "cudaStream_t": ctypes.c_void_p,
}


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes])} for every vca_* prototype in the header."""
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\*|int)\s+(vca_\w+)\s*\(([^)]*)\)\s*;", txt):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes = []
        if args:
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.POINTER(ConvGeom) if "ConvGeom" in a else ctypes.c_void_p)
                else:
                    ty = a.rsplit(" ", 1)[0].replace("const ", "").strip()
                    argtypes.append(_CT[ty])
        protos[name] = (ctypes.c_char_p if ret.startswith("const char") else ctypes.c_int, argtypes)
    return protos


class VcaError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise VcaError(f"{LIB_PATH} is missing -- build it with `python visual-context-attentional-gan_b200/build.py` "
                           "(there is no CPU or torch fallback)")
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        self.launches = 0
        self._prof = None
        for name, (res, args) in self.protos.items():
            fn = getattr(self.cdll, name)
            fn.restype, fn.argtypes = res, args
        # tuning switches from the environment, e.g. VCA_OPTS="wgws_waves=2,hs_mode=0" (see vca_set_option)
        for kv in filter(None, os.environ.get("VCA_OPTS", "").split(",")):
            k, v = kv.split("=")
            if self.cdll.vca_set_option(k.strip().encode(), int(v)) != 0:
                raise VcaError(f"VCA_OPTS: {self.cdll.vca_last_error().decode()}")

    def call(self, name, *args):
        fn = getattr(self.cdll, name)
        conv = []
        for a in args:
            if isinstance(a, torch.Tensor):
                conv.append(ctypes.c_void_p(a.data_ptr()))
            elif isinstance(a, ConvGeom):
                conv.append(ctypes.byref(a))
            else:
                conv.append(a)
        conv.append(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        prof = self._prof
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        rc = fn(*conv)
        self.launches += 1
        if rc != 0:
            raise VcaError(f"{name} failed ({rc}): {self.cdll.vca_last_error().decode()}")
        if prof is not None:
            e1.record()
            flops, tag = 0, ""
            for a in args:
                if isinstance(a, ConvGeom):
                    flops = 2 * a.N * a.OD * a.OH * a.OW * a.Cout * a.Cin * a.KD * a.KH * a.KW
                    if (a.KH, a.KW, a.ph, a.pw) == (5, 3, 2, 1):
                        # pixel-pair merged 5x5 conv (ops._conv5_via_pairs; the model has no native 5x3 conv): count the
                        # ALGORITHMIC work of the original (C/2 -> C/2, 5x5) layer, not the 1.2x Toeplitz-padded MACs
                        flops = flops * 5 // 6
                    tag = "x".join(str(v) for v in a.key())
            if name == "vca_gemm_simt":
                flops = 2 * args[7] * args[8] * args[9] * args[10]
            if not tag:   # non-conv entry points: the integer arguments (shapes, modes) identify the call
                tag = "x".join(str(a) for a in args if isinstance(a, int) and not isinstance(a, bool))
            prof.append((name, tag, flops, e0, e1))

    def try_call(self, name, *args):
        """Like call(), but returns False (instead of raising) when the entry point reports VCA_ERR_UNSUPPORTED (-3)."""
        try:
            self.call(name, *args)
            return True
        except VcaError as e:
            if "(-3)" in str(e):
                return False
            raise

    def profile_step(self, fn):
        """Run fn() with every library call bracketed by CUDA events on the launching stream.
        -> {entry point: {n, ms, flops, top: [(geometry, ms, TFLOP/s)]}}"""
        self._prof = []
        try:
            fn()
            torch.cuda.synchronize()
            rows = [(n, t, f, a.elapsed_time(b)) for n, t, f, a, b in self._prof]
        finally:
            self._prof = None
        agg = {}
        for n, t, f, ms in rows:
            d = agg.setdefault(n, {"n": 0, "ms": 0.0, "flops": 0, "by_geom": {}})
            d["n"] += 1; d["ms"] += ms; d["flops"] += f
            if t:
                g = d["by_geom"].setdefault(t, [0, 0.0, 0])
                g[0] += 1; g[1] += ms; g[2] += f
        for d in agg.values():
            top = sorted(d.pop("by_geom").items(), key=lambda kv: -kv[1][1])[:200]
            d["top"] = [(k, v[0], round(v[1], 3), round(v[2] / max(v[1], 1e-9) / 1e9, 2)) for k, v in top]
        return agg

    def query(self, name, *args):
        fn = getattr(self.cdll, name)
        return fn(*[ctypes.byref(a) if isinstance(a, ConvGeom) else a for a in args])


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
