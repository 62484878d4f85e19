"""ctypes binding of the CPU oracle (oracle/pbs_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.  Parity status: pinned on the reference's goldens down to
decryption; unpinned only at the PBS-output-word level (see pbs_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpbs_oracle.so")
_LIB_PATH_AVX512 = os.path.join(_HERE, "libpbs_oracle_avx512.so")


def _host_has_avx512() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = set(line.split(":", 1)[1].split())
                    return {"avx512f", "avx512dq", "avx512vl", "avx512bw"} <= flags
    except OSError:
        pass
    return False


def lib_path() -> str:
    """The build the loader picks: the AVX-512 twin when the host has it
    (ORACLE_ISA=avx2 forces the baseline build)."""
    if os.environ.get("ORACLE_ISA", "") != "avx2" and os.path.exists(_LIB_PATH_AVX512) and _host_has_avx512():
        return _LIB_PATH_AVX512
    return _LIB_PATH


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("pbs_oracle.c", "pbs_oracle.h", "tfhe_csprng.c", "tfhe_csprng.h")]
    stale = (
        force
        or not os.path.exists(_LIB_PATH)
        or not os.path.exists(_LIB_PATH_AVX512)
        or min(os.path.getmtime(_LIB_PATH), os.path.getmtime(_LIB_PATH_AVX512)) < max(os.path.getmtime(f) for f in srcs)
    )
    if stale:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "all"], env=env)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(lib_path())
        _declare(_lib)
    return _lib


class _Rng(C.Structure):
    _fields_ = [("s", C.c_uint64 * 4)]


class _PbsParams(C.Structure):
    _fields_ = [
        ("n", C.c_uint32),
        ("k", C.c_uint32),
        ("N", C.c_uint32),
        ("base_log", C.c_uint32),
        ("level_count", C.c_uint32),
        ("grouping_factor", C.c_uint32),
        ("centered_ms", C.c_int),
        ("num_many_lut", C.c_uint32),
        ("lut_stride", C.c_uint32),
    ]


_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_u32p = C.POINTER(C.c_uint32)
_f64p = C.POINTER(C.c_double)


def _declare(l):
    u32, u64, i32, vp = C.c_uint32, C.c_uint64, C.c_int32, C.c_void_p
    rp = C.POINTER(_Rng)
    sig = {
        "orc_rng_seed": (None, [rp, u64]),
        "orc_rng_next": (u64, [rp]),
        "orc_fill_uniform": (None, [rp, _u64p, C.c_size_t]),
        "orc_fill_binary": (None, [rp, _u64p, C.c_size_t]),
        "orc_tuniform": (C.c_int64, [rp, u32]),
        "orc_modulus_switch": (u64, [u64, u32]),
        "orc_decomposer_init_state": (u64, [u64, u32, u32]),
        "orc_decompose": (None, [u64, u32, u32, _i64p]),
        "orc_closest_representable": (u64, [u64, u32, u32]),
        "orc_monomial_div": (None, [_u64p, _u64p, u32, u32]),
        "orc_monomial_mul_and_subtract": (None, [_u64p, _u64p, u32, u32]),
        "orc_negacyclic_mul_add_exact": (None, [_u64p, _i64p, _u64p, u32]),
        "orc_lwe_encrypt": (None, [rp, _u64p, u32, u64, i32, _u64p]),
        "orc_lwe_decrypt": (u64, [_u64p, u32, _u64p]),
        "orc_glwe_encrypt_assign": (None, [rp, _u64p, u32, u32, i32, _u64p, _u64p]),
        "orc_gen_bsk": (None, [rp, _u64p, u32, _u64p, u32, u32, u32, u32, i32, _u64p]),
        "orc_gen_multi_bit_bsk": (None, [rp, _u64p, u32, _u64p, u32, u32, u32, u32, u32, i32, _u64p]),
        "orc_gen_ksk": (None, [rp, _u64p, u32, _u64p, u32, u32, u32, i32, _u64p]),
        "orc_keyswitch": (None, [_u64p, u32, u32, u32, u32, _u64p, _u64p]),
        "orc_keyswitch_batch": (None, [_u64p, u32, u32, u32, u32, _u64p, _u64p, u32, u32]),
        "orc_lwe_modulus_switch": (None, [_u64p, u32, u32, C.c_int, _u32p]),
        "orc_centered_ms_body_correction": (u64, [_u64p, u32, u32]),
        "orc_make_lut": (None, [_u64p, u32, u64, u32, u32, _u64p]),
        "orc_sample_extract": (None, [_u64p, u32, u32, u32, _u64p]),
        "orc_fft_plan_new": (vp, [u32]),
        "orc_fft_plan_free": (None, [vp]),
        "orc_fft_forward_integer": (None, [vp, _i64p, _f64p, _f64p]),
        "orc_fft_forward_real": (None, [vp, _f64p, _f64p, _f64p]),
        "orc_fft_forward_torus": (None, [vp, _u64p, _f64p, _f64p]),
        "orc_fft_add_backward_torus": (None, [vp, _f64p, _f64p, _u64p]),
        "orc_bsk_to_fourier": (None, [vp, _u64p, C.c_size_t, _f64p, _f64p]),
        "orc_blind_rotate_fft": (None, [vp, _u64p, _u32p, _f64p, _f64p, u32, u32, u32, u32, u32]),
        "orc_blind_rotate_exact": (None, [_u64p, _u32p, _u64p, u32, u32, u32, u32, u32]),
        "orc_add_external_product_fft": (None, [vp, _u64p, _f64p, _f64p, _u64p, u32, u32, u32, u32]),
        "orc_add_external_product_exact": (None, [_u64p, _u64p, _u64p, u32, u32, u32, u32]),
        "orc_multi_bit_blind_rotate_fft": (None, [vp, _u64p, _u64p, _f64p, _f64p, u32, u32, u32, u32, u32, u32]),
        "orc_multi_bit_blind_rotate_exact": (None, [_u64p, _u64p, _u64p, u32, u32, u32, u32, u32, u32]),
        "orc_pbs_batch": (
            None,
            [vp, C.POINTER(_PbsParams), _u64p, _f64p, _f64p, _u64p, _u64p, _u64p, _u64p, _u64p, _u64p, u32, C.c_int, u32],
        ),
        "orc_max_threads": (u32, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(l, name)
        fn.restype = res
        fn.argtypes = args


def _p(a, typ):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be contiguous"
    return a.ctypes.data_as(typ)


def u64p(a):
    assert a is None or a.dtype == np.uint64
    return _p(a, _u64p)


def i64p(a):
    assert a is None or a.dtype == np.int64
    return _p(a, _i64p)


def u32p(a):
    assert a is None or a.dtype == np.uint32
    return _p(a, _u32p)


def f64p(a):
    assert a is None or a.dtype == np.float64
    return _p(a, _f64p)


# --------------------------------------------------------------------------
# parameter sets (reference file:line in comments)
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class Params:
    name: str
    n: int  # small LWE dimension
    k: int  # GLWE dimension
    N: int  # polynomial size
    pbs_base_log: int
    pbs_level: int
    ks_base_log: int
    ks_level: int
    lwe_noise_log2: int  # TUniform bound for small-key encryptions / KSK
    glwe_noise_log2: int  # TUniform bound for GLWE / BSK
    message_bits: int = 2
    carry_bits: int = 2
    grouping_factor: int = 1
    centered_ms: bool = True

    @property
    def big_n(self) -> int:
        return self.k * self.N

    @property
    def p(self) -> int:  # message modulus incl. carry
        return 1 << (self.message_bits + self.carry_bits)

    @property
    def delta(self) -> int:  # one padding bit
        return (1 << 63) // self.p

    @property
    def log_2N(self) -> int:
        return self.N.bit_length()  # log2(N) + 1

    @property
    def ggsw_polys(self) -> int:
        return self.pbs_level * (self.k + 1) * (self.k + 1)

    @property
    def num_ggsw(self) -> int:
        if self.grouping_factor > 1:
            return (self.n // self.grouping_factor) << self.grouping_factor
        return self.n


# tfhe/src/shortint/parameters/v1_4/classic/tuniform/p_fail_2_minus_128/ks_pbs.rs:29-47
PARAM_MESSAGE_2_CARRY_2_KS_PBS = Params(
    "PARAM_MESSAGE_2_CARRY_2_KS_PBS_TUNIFORM_2M128",
    n=918, k=1, N=2048, pbs_base_log=23, pbs_level=1, ks_base_log=4, ks_level=4,
    lwe_noise_log2=45, glwe_noise_log2=17,
)
# tfhe/src/shortint/parameters/v1_1/multi_bit/tuniform/p_fail_2_minus_128/ks_pbs.rs:118-137
PARAM_MULTI_BIT_GROUP_3_MESSAGE_2_CARRY_2_KS_PBS = Params(
    "PARAM_MULTI_BIT_GROUP_3_MESSAGE_2_CARRY_2_KS_PBS_TUNIFORM_2M128",
    n=918, k=1, N=2048, pbs_base_log=15, pbs_level=2, ks_base_log=3, ks_level=6,
    lwe_noise_log2=45, glwe_noise_log2=17, grouping_factor=3, centered_ms=False,
)
# small, fast sets used by the CPU tests (not reference sets; noise chosen so
# that decryption is always correct)
# tfhe/src/shortint/parameters/v1_0/classic/tuniform/p_fail_2_minus_128/ks_pbs.rs:11-24
PARAM_MESSAGE_1_CARRY_1_KS_PBS = Params(
    "PARAM_MESSAGE_1_CARRY_1_KS_PBS_TUNIFORM_2M128",
    n=879, k=4, N=512, pbs_base_log=23, pbs_level=1, ks_base_log=5, ks_level=3,
    lwe_noise_log2=46, glwe_noise_log2=17, message_bits=1, carry_bits=1,
)
# ks_pbs.rs:67-92 (the drift-technique modulus switch is a host-side step
# before the PBS; the PBS kernels see the plain modulus switch)
PARAM_MESSAGE_3_CARRY_3_KS_PBS = Params(
    "PARAM_MESSAGE_3_CARRY_3_KS_PBS_TUNIFORM_2M128",
    n=1077, k=1, N=8192, pbs_base_log=15, pbs_level=2, ks_base_log=4, ks_level=5,
    lwe_noise_log2=41, glwe_noise_log2=3, message_bits=3, carry_bits=3,
)
# N = 8192 toy for the register kernel of csrc/pbs_n8192.cuh
TOY_N8192 = Params("TOY_N8192_3_3", n=40, k=1, N=8192, pbs_base_log=15, pbs_level=2, ks_base_log=4, ks_level=5,
                   lwe_noise_log2=40, glwe_noise_log2=3, message_bits=3, carry_bits=3)
# N = 512 toys for the register kernel of csrc/pbs_n512.cuh (k = 1..4, short n)
TOY_N512 = {k: Params("TOY_N512_K%d" % k, n=24, k=k, N=512, pbs_base_log=23, pbs_level=1, ks_base_log=5,
                      ks_level=3, lwe_noise_log2=40, glwe_noise_log2=17, message_bits=1, carry_bits=1)
            for k in (1, 2, 3, 4)}
TOY_K1 = Params("TOY_N256_K1", n=24, k=1, N=256, pbs_base_log=23, pbs_level=1,
                ks_base_log=4, ks_level=4, lwe_noise_log2=30, glwe_noise_log2=10)
TOY_K2_L2 = Params("TOY_N256_K2_L2", n=20, k=2, N=256, pbs_base_log=12, pbs_level=2,
                   ks_base_log=3, ks_level=5, lwe_noise_log2=30, glwe_noise_log2=10,
                   centered_ms=False)
TOY_MB3 = Params("TOY_N256_MB3", n=24, k=1, N=256, pbs_base_log=15, pbs_level=2,
                 ks_base_log=3, ks_level=6, lwe_noise_log2=30, glwe_noise_log2=10,
                 grouping_factor=3, centered_ms=False)


class Rng:
    def __init__(self, seed: int):
        self._r = _Rng()
        lib().orc_rng_seed(C.byref(self._r), C.c_uint64(seed & (2**64 - 1)))

    @property
    def ref(self):
        return C.byref(self._r)

    def next(self) -> int:
        return int(lib().orc_rng_next(self.ref))

    def uniform(self, count: int) -> np.ndarray:
        out = np.empty(count, dtype=np.uint64)
        lib().orc_fill_uniform(self.ref, u64p(out), count)
        return out

    def binary(self, count: int) -> np.ndarray:
        out = np.empty(count, dtype=np.uint64)
        lib().orc_fill_binary(self.ref, u64p(out), count)
        return out


class FftPlan:
    def __init__(self, N: int):
        self.N = N
        self.M = N // 2
        self._h = lib().orc_fft_plan_new(N)

    def __del__(self):
        try:
            if self._h:
                lib().orc_fft_plan_free(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def h(self):
        return self._h

    def forward_integer(self, poly: np.ndarray):
        re, im = np.empty(self.M), np.empty(self.M)
        lib().orc_fft_forward_integer(self._h, i64p(np.ascontiguousarray(poly, dtype=np.int64)), f64p(re), f64p(im))
        return re, im

    def forward_real(self, poly: np.ndarray):
        re, im = np.empty(self.M), np.empty(self.M)
        lib().orc_fft_forward_real(self._h, f64p(np.ascontiguousarray(poly, dtype=np.float64)), f64p(re), f64p(im))
        return re, im

    def forward_torus(self, poly: np.ndarray):
        re, im = np.empty(self.M), np.empty(self.M)
        lib().orc_fft_forward_torus(self._h, u64p(np.ascontiguousarray(poly, dtype=np.uint64)), f64p(re), f64p(im))
        return re, im

    def add_backward_torus(self, re, im, poly_inout: np.ndarray):
        lib().orc_fft_add_backward_torus(
            self._h, f64p(np.ascontiguousarray(re)), f64p(np.ascontiguousarray(im)), u64p(poly_inout)
        )


@dataclass
class KeySet:
    """Secret + evaluation keys for one parameter set, all numpy u64."""

    params: Params
    lwe_sk: np.ndarray  # n
    glwe_sk: np.ndarray  # k*N  (= big LWE key)
    bsk: np.ndarray  # standard domain, [num_ggsw][l][k+1][k+1][N]
    ksk: np.ndarray | None  # [kN][l_ks][n+1]
    _fourier: tuple | None = None
    _plan: FftPlan | None = None

    @property
    def plan(self) -> FftPlan:
        if self._plan is None:
            self._plan = FftPlan(self.params.N)
        return self._plan

    def fourier_bsk(self):
        if self._fourier is None:
            polys = self.bsk.size // self.params.N
            re = np.empty(polys * self.params.N // 2)
            im = np.empty(polys * self.params.N // 2)
            lib().orc_bsk_to_fourier(self.plan.h, u64p(self.bsk), polys, f64p(re), f64p(im))
            self._fourier = (re, im)
        return self._fourier


def keygen(params: Params, seed: int, with_ksk: bool = True) -> KeySet:
    rng = Rng(seed)
    lwe_sk = rng.binary(params.n)
    glwe_sk = rng.binary(params.k * params.N)
    p = params
    bsk = np.zeros(p.num_ggsw * p.ggsw_polys * p.N, dtype=np.uint64)
    if p.grouping_factor > 1:
        lib().orc_gen_multi_bit_bsk(rng.ref, u64p(lwe_sk), p.n, u64p(glwe_sk), p.k, p.N, p.pbs_base_log,
                                    p.pbs_level, p.grouping_factor, p.glwe_noise_log2, u64p(bsk))
    else:
        lib().orc_gen_bsk(rng.ref, u64p(lwe_sk), p.n, u64p(glwe_sk), p.k, p.N, p.pbs_base_log, p.pbs_level,
                          p.glwe_noise_log2, u64p(bsk))
    ksk = None
    if with_ksk:
        ksk = np.zeros(p.big_n * p.ks_level * (p.n + 1), dtype=np.uint64)
        lib().orc_gen_ksk(rng.ref, u64p(glwe_sk), p.big_n, u64p(lwe_sk), p.n, p.ks_base_log, p.ks_level,
                          p.lwe_noise_log2, u64p(ksk))
    return KeySet(params, lwe_sk, glwe_sk, bsk, ksk)


def lwe_encrypt_batch(rng: Rng, key: np.ndarray, plaintexts, noise_log2: int) -> np.ndarray:
    n = key.size
    pts = np.asarray(plaintexts, dtype=np.uint64)
    out = np.empty((pts.size, n + 1), dtype=np.uint64)
    for s in range(pts.size):
        lib().orc_lwe_encrypt(rng.ref, u64p(key), n, C.c_uint64(int(pts[s])), noise_log2, u64p(out[s]))
    return out


def lwe_decrypt_batch(key: np.ndarray, cts: np.ndarray) -> np.ndarray:
    """b - <a, s> for each row (vectorised in numpy, wrapping)."""
    cts = np.ascontiguousarray(cts, dtype=np.uint64).reshape(-1, key.size + 1)
    with np.errstate(over="ignore"):
        dot = (cts[:, :-1] * key[None, :]).sum(axis=1, dtype=np.uint64)
        return cts[:, -1] - dot


def decode(plaintexts: np.ndarray, delta: int, p: int) -> np.ndarray:
    """divide_round(pt, delta) mod p  (algorithms/test/mod.rs:489-491)."""
    pts = np.asarray(plaintexts, dtype=np.uint64)
    half = np.uint64(delta // 2)
    with np.errstate(over="ignore"):
        return ((pts + half) // np.uint64(delta)) % np.uint64(p)


def make_lut(params: Params, f_values, delta: int | None = None, p: int | None = None) -> np.ndarray:
    p = p or params.p
    delta = delta if delta is not None else params.delta
    fv = np.asarray([int(v) % (1 << 64) for v in f_values], dtype=np.uint64)
    assert fv.size == p
    out = np.empty((params.k + 1) * params.N, dtype=np.uint64)
    lib().orc_make_lut(u64p(fv), p, C.c_uint64(delta), params.k, params.N, u64p(out))
    return out


def keyswitch_batch(keys: KeySet, cts_in: np.ndarray, threads: int = 0) -> np.ndarray:
    p = keys.params
    cts_in = np.ascontiguousarray(cts_in, dtype=np.uint64).reshape(-1, p.big_n + 1)
    out = np.empty((cts_in.shape[0], p.n + 1), dtype=np.uint64)
    lib().orc_keyswitch_batch(u64p(keys.ksk), p.big_n, p.n, p.ks_base_log, p.ks_level, u64p(cts_in), u64p(out),
                              cts_in.shape[0], threads)
    return out


def modulus_switch_lwe(ct: np.ndarray, log_modulus: int, centered: bool) -> np.ndarray:
    ct = np.ascontiguousarray(ct, dtype=np.uint64)
    out = np.empty(ct.size, dtype=np.uint32)
    lib().orc_lwe_modulus_switch(u64p(ct), ct.size - 1, log_modulus, int(centered), u32p(out))
    return out


def sample_extract(glwe: np.ndarray, k: int, N: int, nth: int) -> np.ndarray:
    """extract_lwe_sample_from_glwe_ciphertext (glwe_sample_extraction.rs:119-165)."""
    glwe = np.ascontiguousarray(glwe, dtype=np.uint64)
    out = np.empty(k * N + 1, dtype=np.uint64)
    lib().orc_sample_extract(u64p(glwe), k, N, nth, u64p(out))
    return out


def pbs_batch(keys: KeySet, luts: np.ndarray, cts_in: np.ndarray, *, lut_idx=None, in_idx=None, out_idx=None,
              exact: bool = False, threads: int = 0, num_many_lut: int = 1, lut_stride: int = 0,
              centered_ms: bool | None = None, count: int | None = None, out_rows: int | None = None) -> np.ndarray:
    """Batched PBS with the C-ABI's indexing conventions (see pbs_oracle.h)."""
    p = keys.params
    cts_in = np.ascontiguousarray(cts_in, dtype=np.uint64).reshape(-1, p.n + 1)
    luts = np.ascontiguousarray(luts, dtype=np.uint64)
    count = cts_in.shape[0] if count is None else count
    rows = out_rows if out_rows is not None else count
    out = np.zeros((num_many_lut * rows, p.big_n + 1), dtype=np.uint64)
    prm = _PbsParams(p.n, p.k, p.N, p.pbs_base_log, p.pbs_level, p.grouping_factor,
                     int(p.centered_ms if centered_ms is None else centered_ms), num_many_lut, lut_stride)
    cv = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.uint64)
    lut_idx, in_idx, out_idx = cv(lut_idx), cv(in_idx), cv(out_idx)
    if exact:
        bre = bim = None
    else:
        bre, bim = keys.fourier_bsk()
    assert rows == count or out_idx is not None
    lib().orc_pbs_batch(keys.plan.h, C.byref(prm), u64p(keys.bsk), f64p(bre), f64p(bim), u64p(luts), u64p(lut_idx),
                        u64p(cts_in), u64p(in_idx), u64p(out), u64p(out_idx), count, int(exact), threads)
    return out


def max_threads() -> int:
    """Host threads the oracle may really use: min(CPU affinity mask, cgroup
    cpu quota) -- a container often sees all the host's cores through nproc but
    is only allowed a share of them.  OMP_NUM_THREADS is deliberately NOT
    consulted: torchrun exports OMP_NUM_THREADS=1 to every rank, which would
    turn the "all host cores" CPU arm into a single-thread run; callers pass
    the count explicitly (`threads=`) to every batched entry point."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, q // per))
        except Exception:
            pass
    return max(1, n)
