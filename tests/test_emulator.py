"""CPU tests of the kernels' device code through the CTA emulator (tests/emu):
the same B200_HD phase functions the sm_100a kernels inline are replayed
thread by thread and compared with the oracle.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def emu():
    from tests.emu.build_emu import build

    return C.CDLL(build())


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _bitrev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2)


def test_register_fft_matches_oracle(oracle, emu):
    N, M = 2048, 1024
    rng = np.random.default_rng(1)
    poly = rng.integers(-(1 << 22), 1 << 22, size=N).astype(np.int64)
    re, im = oracle.FftPlan(N).forward_integer(poly)
    X = re + 1j * im  # natural order, X[k] = p(t^(1-4k))
    z = np.empty(2 * M)
    z[0::2], z[1::2] = poly[:M], poly[M:]
    out = np.empty(2 * M)
    emu.emu_fft1024_fwd(_vp(z), _vp(out))
    Y = out[0::2] + 1j * out[1::2]  # slot order: Y[pos] = p(t^(1+4 bitrev(pos)))
    perm = np.array([(-_bitrev(p, 10)) % M for p in range(M)])
    assert np.abs(Y - X[perm]).max() < 1e-14 * np.abs(X).max() * 64
    back = np.empty(2 * M)
    emu.emu_fft1024_inv(_vp(out), _vp(back))
    assert np.abs(back / 1024 - z).max() < 1e-6


def test_register_fft256_matches_oracle(oracle, emu):
    """The 16 x 16 transform of the N = 512 kernel (one exchange, 16 threads per
    polynomial) against the oracle's FFT: slot pos holds p(t^(1+4 bitrev8(pos)))."""
    N, M = 512, 256
    rng = np.random.default_rng(5)
    poly = rng.integers(-(1 << 22), 1 << 22, size=N).astype(np.int64)
    re, im = oracle.FftPlan(N).forward_integer(poly)
    X = re + 1j * im
    z = np.empty(2 * M)
    z[0::2], z[1::2] = poly[:M], poly[M:]
    out = np.empty(2 * M)
    emu.emu_fft256_fwd(_vp(z), _vp(out))
    Y = out[0::2] + 1j * out[1::2]
    perm = np.array([(-_bitrev(p, 8)) % M for p in range(M)])
    assert np.abs(Y - X[perm]).max() < 1e-14 * np.abs(X).max() * 64
    back = np.empty(2 * M)
    emu.emu_fft256_inv(_vp(out), _vp(back))
    assert np.abs(back / 256 - z).max() < 1e-7


def test_register_fft4096_matches_oracle(oracle, emu):
    """The 16 x 16 x 16 transform of the N = 8192 kernel (256 threads per polynomial) against the oracle's FFT."""
    N, M = 8192, 4096
    rng = np.random.default_rng(9)
    poly = rng.integers(-(1 << 14), 1 << 14, size=N).astype(np.int64)
    re, im = oracle.FftPlan(N).forward_integer(poly)
    X = re + 1j * im
    z = np.empty(2 * M)
    z[0::2], z[1::2] = poly[:M], poly[M:]
    out = np.empty(2 * M)
    emu.emu_fft4096_fwd(_vp(z), _vp(out))
    Y = out[0::2] + 1j * out[1::2]
    perm = np.array([(-_bitrev(p, 12)) % M for p in range(M)])
    assert np.abs(Y - X[perm]).max() < 1e-14 * np.abs(X).max() * 256
    back = np.empty(2 * M)
    emu.emu_fft4096_inv(_vp(out), _vp(back))
    assert np.abs(back / 4096 - z).max() < 1e-6


def test_tensor_memory_exchange_variant_is_bit_identical(emu):
    """Exchange 2 through the tensor-memory model (tmem_x2.cuh) with the matching
    exchange-1 layout moves the same values to the same registers: both
    transforms agree bit for bit with the shared-memory variant."""
    M = 1024
    rng = np.random.default_rng(7)
    z = rng.integers(-(1 << 22), 1 << 22, size=2 * M).astype(np.float64)
    a, b = np.empty(2 * M), np.empty(2 * M)
    emu.emu_fft1024_fwd(_vp(z), _vp(a))
    emu.emu_fft1024_fwd_tmem(_vp(z), _vp(b))
    assert np.array_equal(a, b)
    c, d = np.empty(2 * M), np.empty(2 * M)
    emu.emu_fft1024_inv(_vp(a), _vp(c))
    emu.emu_fft1024_inv_tmem(_vp(a), _vp(d))
    assert np.array_equal(c, d)


def _p22(oracle, n):
    return oracle.Params("P22_n%d" % n, n=n, k=1, N=2048, pbs_base_log=23, pbs_level=1, ks_base_log=4, ks_level=4,
                         lwe_noise_log2=45, glwe_noise_log2=17)


def _emu_pbs(emu, keys, lut, cts, centered, many=1, stride=0, variant=1):
    P = keys.params
    bskf = np.empty(P.n * 4 * 1024 * 2)
    emu.emu_bsk_convert_p22(_vp(keys.bsk), P.n, _vp(bskf))
    out = np.zeros((many, len(cts), 2049), dtype=np.uint64)
    for s in range(len(cts)):
        fn = {1: emu.emu_pbs_p22, 3: emu.emu_pbs_p22_v3}[variant]
        fn(_vp(bskf), _vp(lut), _vp(cts[s]), P.n, P.pbs_base_log, int(centered), many, stride, len(cts),
           _vp(out[0, s]))
    return out


def test_exchange_layouts_are_bank_conflict_free(emu):
    assert emu.emu_exchange_conflict_audit() == 1


@pytest.mark.parametrize("variant", [1, 3])
def test_emulated_kernel_single_cmux_word_level(oracle, keyset, emu, variant):
    """n = 1: one external product; word-level agreement with the exact oracle
    within the f64 FFT noise floor."""
    P = _p22(oracle, 1)
    keys = keyset(P, seed=11, with_ksk=False)
    msgs = np.arange(8) % 16
    cts = oracle.lwe_encrypt_batch(oracle.Rng(3), keys.lwe_sk, msgs.astype(np.uint64) * np.uint64(P.delta), 45)
    lut = oracle.make_lut(P, [(5 * i + 3) % 16 for i in range(16)])
    out = _emu_pbs(emu, keys, lut, cts, True, variant=variant)[0]
    ref = oracle.pbs_batch(keys, lut, cts, exact=True)
    assert np.abs((out - ref).astype(np.int64)).max() < (1 << 43)


@pytest.mark.parametrize("variant", [1, 3])
@pytest.mark.parametrize("centered", [True, False])
def test_emulated_kernel_decrypts(oracle, keyset, emu, centered, variant):
    P = _p22(oracle, 16)
    keys = keyset(P, seed=7, with_ksk=False)
    msgs = np.arange(8) % 16
    cts = oracle.lwe_encrypt_batch(oracle.Rng(5), keys.lwe_sk, msgs.astype(np.uint64) * np.uint64(P.delta), 45)
    f = [(5 * i + 3) % 16 for i in range(16)]
    lut = oracle.make_lut(P, f)
    out = _emu_pbs(emu, keys, lut, cts, centered, variant=variant)[0]
    dec = oracle.decode(oracle.lwe_decrypt_batch(keys.glwe_sk, out), P.delta, 16)
    assert np.array_equal(dec, np.array([f[m] for m in msgs]))
    ref = oracle.pbs_batch(keys, lut, cts, centered_ms=centered)
    refdec = oracle.decode(oracle.lwe_decrypt_batch(keys.glwe_sk, ref), P.delta, 16)
    assert np.array_equal(dec, refdec)


@pytest.mark.parametrize("variant", [1, 3])
def test_emulated_kernel_zero_mask_is_bit_exact(oracle, keyset, emu, variant):
    """All-zero mask: every CMUX is skipped, the path is integer only (LUT
    rotation by b_hat + sample extract) and must be bit-identical."""
    P = _p22(oracle, 4)
    keys = keyset(P, seed=7, with_ksk=False)
    lut = oracle.make_lut(P, list(range(16)))
    cts = np.zeros((5, P.n + 1), dtype=np.uint64)
    cts[:, -1] = np.array([0, 1 << 59, 3 << 59, (1 << 63) + (5 << 59), (1 << 64) - 1], dtype=np.uint64)
    for centered in (False, True):
        out = _emu_pbs(emu, keys, lut, cts, centered, many=2, stride=3, variant=variant)
        ref = oracle.pbs_batch(keys, lut, cts, centered_ms=centered, num_many_lut=2, lut_stride=3)
        assert np.array_equal(out.reshape(-1, 2049), ref)


@pytest.mark.parametrize("grouping,level,base_log", [(3, 2, 15), (2, 2, 15), (3, 1, 23), (4, 1, 22)])
def test_emulated_multibit_kernel(oracle, keyset, emu, grouping, level, base_log):
    """Fast multi-bit kernel (N=2048, k=1): decrypt-equal to the oracle on
    random inputs and on the zero-mask path with many-LUT outputs."""
    P = oracle.Params(f"MB_g{grouping}_l{level}", n=12, k=1, N=2048, pbs_base_log=base_log, pbs_level=level,
                      ks_base_log=3, ks_level=6, lwe_noise_log2=45, glwe_noise_log2=17, grouping_factor=grouping,
                      centered_ms=False)
    keys = keyset(P, seed=19, with_ksk=False)
    bskf = np.empty(P.num_ggsw * P.ggsw_polys * 1024 * 2)
    emu.emu_bsk_convert_mb(_vp(keys.bsk), P.n, level, grouping, _vp(bskf))
    msgs = np.arange(6) % 16
    cts = oracle.lwe_encrypt_batch(oracle.Rng(5), keys.lwe_sk, msgs.astype(np.uint64) * np.uint64(P.delta), 45)
    f = [(7 * i + 2) % 16 for i in range(16)]
    lut = oracle.make_lut(P, f)

    def run(cts_, many=1, stride=0):
        out = np.zeros((many, len(cts_), 2049), dtype=np.uint64)
        for s in range(len(cts_)):
            emu.emu_pbs_mb(_vp(bskf), _vp(lut), _vp(cts_[s]), P.n, base_log, level, grouping, many, stride,
                           len(cts_), _vp(out[0, s]))
        return out

    out = run(cts)[0]
    dec = oracle.decode(oracle.lwe_decrypt_batch(keys.glwe_sk, out), P.delta, 16)
    assert np.array_equal(dec, np.array([f[m] for m in msgs]))
    ref = oracle.pbs_batch(keys, lut, cts)
    assert np.array_equal(dec, oracle.decode(oracle.lwe_decrypt_batch(keys.glwe_sk, ref), P.delta, 16))
    zero = np.zeros((3, P.n + 1), dtype=np.uint64)
    zero[:, -1] = np.array([0, 3 << 59, (1 << 63) + (5 << 59)], dtype=np.uint64)
    got = run(zero, many=2, stride=3)
    want = oracle.pbs_batch(keys, lut, zero, num_many_lut=2, lut_stride=3)
    assert np.array_equal(oracle.decode(oracle.lwe_decrypt_batch(keys.glwe_sk, got.reshape(-1, 2049)), P.delta, 16),
                          oracle.decode(oracle.lwe_decrypt_batch(keys.glwe_sk, want), P.delta, 16))


@pytest.mark.parametrize("base_log,level", [(15, 2), (22, 1), (23, 1), (7, 2)])
def test_digits_u32_matches_reference_decomposer_except_ties(oracle, emu, base_log, level):
    """The multi-bit register kernels decompose the 32-bit accumulator word.
    For every non-tie input the digits are those of the reference decomposer
    (decomposer.rs:163-188, iter.rs:131-151) on the word placed in the top half
    of a u64; on an exact tie of the dropped bits the closest representable is
    the EVEN one (the reference rounds up; see digits_u32 for why), and the
    digits still recompose to it."""
    R = base_log * level
    drop = 32 - R
    rng = np.random.default_rng(5)
    xs = rng.integers(0, 1 << 32, size=4000, dtype=np.uint64).astype(np.uint32)
    # force a share of exact ties and near-ties
    xs[:1000] = (xs[:1000] & ~np.uint32((1 << drop) - 1)) | np.uint32(1 << (drop - 1))
    xs[1000:1200] += np.uint32(1)
    out = (C.c_int32 * 2)()
    ref = (C.c_int64 * level)()
    ties = 0
    for x in xs:
        x = int(x)
        emu.emu_digits_u32(x, base_log, level, out)
        got = [out[i] for i in range(level)]
        low, half, q = x & ((1 << drop) - 1), 1 << (drop - 1), x >> drop
        if low == half:
            ties += 1
            want_q = q + (q & 1)  # to even
        else:
            oracle.lib().orc_decompose(C.c_uint64(x << 32), base_log, level, ref)
            assert got == [ref[i] for i in range(level)], hex(x)
            want_q = q + (1 if low > half else 0)
        assert all(-(1 << (base_log - 1)) <= d <= (1 << (base_log - 1)) for d in got)
        # digit[0] is level l (weight 1 in units of 2^drop), digit[i] weight B^i
        recomposed = sum(d << (base_log * i) for i, d in enumerate(got))
        assert (recomposed - want_q) % (1 << R) == 0, hex(x)
        # reference-exact switch (b200_set_multibit_tie_rule): the reference decomposer on EVERY input, ties included
        emu.emu_digits_u32_reference_ties(x, base_log, level, out)
        oracle.lib().orc_decompose(C.c_uint64(x << 32), base_log, level, ref)
        assert [out[i] for i in range(level)] == [ref[i] for i in range(level)], hex(x)
    assert ties >= 1000


def test_multibit_register_kernel_noise_not_above_oracle(oracle, keyset, emu):
    """Noise regression guard for the 32-bit accumulator of the multi-bit
    register kernels (l = 2, B = 2^15: the case where always-up tie rounding in
    the decomposition gave 5-8x the noise of the u64 reference path): over 8
    steps the emulated kernel's output-noise variance must stay within 2.5x of
    the oracle's FFT mode on the same keys and inputs (it is ~0.5x with ties to
    even, was ~5.6x before)."""
    P = oracle.Params("MB_noise", n=24, k=1, N=2048, pbs_base_log=15, pbs_level=2, ks_base_log=3, ks_level=6,
                      lwe_noise_log2=45, glwe_noise_log2=17, grouping_factor=3, centered_ms=False)
    keys = keyset(P, seed=19, with_ksk=False)
    bskf = np.empty(P.num_ggsw * P.ggsw_polys * 1024 * 2)
    emu.emu_bsk_convert_mb(_vp(keys.bsk), P.n, 2, 3, _vp(bskf))
    ns = 48
    msgs = np.arange(ns) % 16
    cts = oracle.lwe_encrypt_batch(oracle.Rng(5), keys.lwe_sk, msgs.astype(np.uint64) * np.uint64(P.delta), 45)
    lut = oracle.make_lut(P, list(range(16)))
    out = np.zeros((ns, 2049), dtype=np.uint64)
    for s in range(ns):
        emu.emu_pbs_mb(_vp(bskf), _vp(lut), _vp(cts[s]), P.n, 15, 2, 3, 1, 0, ns, _vp(out[s]))
    ref = oracle.pbs_batch(keys, lut, cts)
    expected = msgs.astype(np.uint64) * np.uint64(P.delta)

    def var(o):
        with np.errstate(over="ignore"):
            return ((oracle.lwe_decrypt_batch(keys.glwe_sk, o) - expected).astype(np.int64) / 2.0 ** 64).var()

    assert np.array_equal(oracle.decode(oracle.lwe_decrypt_batch(keys.glwe_sk, out), P.delta, 16), msgs)
    assert var(out) < 2.5 * var(ref), (var(out), var(ref))


@pytest.mark.parametrize("base_log", [15, 14, 12, 9, 4])
def test_two_level_digits_specialisation_is_word_exact(emu, base_log):
    """digits2_u32 (the l = 2 decomposition of csrc/pbs_n8192.cuh) yields the digits of digits_u32<2> on 2^24 walked
    words plus the edge words (top-level balance at +-2^(R-1), digit boundaries at B/2, every pattern of the dropped
    bits), under both tie rules."""
    emu.emu_digits2_mismatches.restype = C.c_uint64
    emu.emu_digits2_mismatches.argtypes = [C.c_uint32, C.c_uint64, C.c_int]
    for ties_even in (1, 0):
        assert emu.emu_digits2_mismatches(base_log, 1 << 24, ties_even) == 0


@pytest.mark.parametrize("centered", [True, False])
def test_n8192_kernel_replay_matches_oracle(oracle, emu, centered):
    """The (N = 8192, k = 1, l = 2) tensor-memory kernel (csrc/pbs_n8192.cuh) replayed on the CPU from its own phase
    functions -- key conversion into the kernel's layout, rotate + decompose once per polynomial with the level-1
    digits packed, 16 x 16 x 16 transforms, MAC in the kernel's (level slot, row) order, u32 accumulators, many-LUT
    sample extract: decrypts like the oracle on the same keys, output PHASES within 2^-20 of the oracle's, and
    zero-mask inputs equal the oracle's words rounded to 32 bits."""
    P = oracle.Params("EMU_N8192", n=12, k=1, N=8192, pbs_base_log=15, pbs_level=2, ks_base_log=4, ks_level=5,
                      lwe_noise_log2=40, glwe_noise_log2=3, message_bits=3, carry_bits=3, centered_ms=centered)
    keys = oracle.keygen(P, 0x8192, with_ksk=False)
    count = 3
    msgs = np.array([5, 17, 30])
    cts = oracle.lwe_encrypt_batch(oracle.Rng(4), keys.lwe_sk, msgs.astype(np.uint64) * np.uint64(P.delta),
                                   P.lwe_noise_log2)
    f = [(5 * i + 3) % P.p for i in range(P.p)]
    lut = oracle.make_lut(P, f)
    bsk = np.empty(P.n * 8 * 4096 * 2)
    emu.emu_bsk_convert_n8192(_vp(np.ascontiguousarray(keys.bsk)), P.n, _vp(bsk))

    def run(c, many=1, stride=0):
        out = np.zeros((many * count, P.N + 1), dtype=np.uint64)
        emu.emu_pbs_n8192(_vp(bsk), _vp(np.ascontiguousarray(lut)), _vp(np.ascontiguousarray(c)), P.n, P.pbs_base_log,
                          int(centered), 1, many, stride, count, _vp(out))
        return out

    got = run(cts)
    ref = oracle.pbs_batch(keys, lut, cts)
    want = np.array([f[m] for m in msgs])
    assert np.array_equal(oracle.decode(oracle.lwe_decrypt_batch(keys.glwe_sk, got), P.delta, P.p), want)
    assert np.array_equal(oracle.decode(oracle.lwe_decrypt_batch(keys.glwe_sk, ref), P.delta, P.p), want)
    d = (oracle.lwe_decrypt_batch(keys.glwe_sk, got) - oracle.lwe_decrypt_batch(keys.glwe_sk, ref)).astype(np.int64)
    assert np.abs(d.astype(np.float64)).max() < 2.0 ** 44
    zero = cts.copy()
    zero[:, :P.n] = 0
    got0 = run(zero, many=2, stride=5)
    ref0 = oracle.pbs_batch(keys, lut, zero, num_many_lut=2, lut_stride=5)
    assert np.array_equal(got0, (ref0 + np.uint64(1 << 31)) & np.uint64(0xFFFFFFFF00000000))
