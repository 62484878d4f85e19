"""CPU tests of the oracle itself (no GPU): every golden vector / doc-test
example the reference holds for the routines restated in oracle/pbs_oracle.c,
plus the reference's semantic assertions (decrypt(PBS(Enc m)) == f(m))."""
import ctypes as C
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
U64 = (1 << 64) - 1


def _fft16_reference_input(n=2048):
    # closed-form input of tfhe/src/core_crypto/gpu/algorithms/test/fft/mod.rs:51-71
    def coeff(k):
        bits = (k * 0x517CC1B727220A95) & U64
        bits = ((bits << 17) | (bits >> 47)) & U64
        bits ^= 0xDEADBEEFCAFEBABE
        if bits >= 1 << 63:
            bits -= 1 << 64
        return float(bits) / float((1 << 63) - 1)

    return np.array([coeff(k) for k in range(n)])


def test_forward_fft_matches_reference_golden(oracle):
    """KAT: the reference's committed H100 spectrum (fft16x4x16_golden_v1.rs)."""
    g = np.load(os.path.join(GOLDEN, "fft16x4x16_golden_v1.npz"))
    want = g["expected_re_bits"].view(np.float64) + 1j * g["expected_im_bits"].view(np.float64)
    plan = oracle.FftPlan(2048)
    re, im = plan.forward_real(_fft16_reference_input())
    err = np.abs((re + 1j * im) - want).max()
    assert err < 1e-12 * np.abs(want).max(), err


def test_fft_roundtrip_error_bound(oracle):
    """fft/tests.rs:38-42: |roundtrip - input| < 2^14 on u64 for every size."""
    rng = oracle.Rng(7)
    for logn in range(6, 13):
        n = 1 << logn
        plan = oracle.FftPlan(n)
        poly = rng.uniform(n)
        re, im = plan.forward_torus(poly)
        back = np.zeros(n, dtype=np.uint64)
        plan.add_backward_torus(re, im, back)
        diff = (back - poly).astype(np.int64)
        assert np.abs(diff).max() < (1 << 14), (n, np.abs(diff).max())


def test_fft_product_vs_naive_convolution(oracle):
    """fft/tests.rs:72-208: FFT product == naive negacyclic product."""
    L = oracle.lib()
    rng = oracle.Rng(11)
    for n in (32, 256, 2048):
        plan = oracle.FftPlan(n)
        a = (rng.uniform(n) % np.uint64(1 << 16)).astype(np.int64) - (1 << 15)
        b = rng.uniform(n)
        exact = np.zeros(n, dtype=np.uint64)
        L.orc_negacyclic_mul_add_exact(oracle.u64p(exact), oracle.i64p(a), oracle.u64p(b), n)
        ar, ai = plan.forward_integer(a)
        br, bi = plan.forward_torus(b)
        pr, pi_ = ar * br - ai * bi, ar * bi + ai * br
        got = np.zeros(n, dtype=np.uint64)
        plan.add_backward_torus(pr, pi_, got)
        diff = (got - exact).astype(np.int64)
        assert np.abs(diff).max() < (1 << 40), (n, np.abs(diff).max())


def test_decomposer_doc_examples(oracle):
    """decomposer.rs doc-tests: closest_representable / decompose bounds."""
    L = oracle.lib()
    # SignedDecomposer::<u32>::new(BaseLog(4), Level(3)).closest_representable(1_340_987_234) == 1_341_128_704
    # restated on the u64 torus by shifting the u32 example up 32 bits
    got = L.orc_closest_representable(C.c_uint64(1_340_987_234 << 32), 4, 3)
    assert got == 1_341_128_704 << 32
    digits = np.zeros(3, dtype=np.int64)
    L.orc_decompose(C.c_uint64(2147483647 << 32), 4, 3, oracle.i64p(digits))
    assert all(-8 <= d <= 8 for d in digits)
    # recomposition: sum digit_level * q / B^level == closest representable
    for x in (0, 1, U64, 0x8000000000000000, 0x123456789ABCDEF0, 0x7FFFFFFFFFFFFFFF):
        for (b, l) in ((4, 4), (23, 1), (15, 2), (3, 6), (8, 3)):
            d = np.zeros(l, dtype=np.int64)
            L.orc_decompose(C.c_uint64(x), b, l, oracle.i64p(d))
            assert all(-(1 << (b - 1)) <= int(v) <= (1 << (b - 1)) for v in d)
            rec = 0
            for t in range(l):
                level = l - t
                rec = (rec + int(d[t]) * (1 << (64 - b * level))) & U64
            assert rec == L.orc_closest_representable(C.c_uint64(x), b, l), (hex(x), b, l)


def test_monomial_doc_examples(oracle):
    """polynomial_algorithms.rs:535-543 (div) and the mul twin."""
    L = oracle.lib()
    inp = np.array([1, 2, 3], dtype=np.uint64)
    out = np.zeros(3, dtype=np.uint64)
    L.orc_monomial_div(oracle.u64p(out), oracle.u64p(inp), 3, 2)
    # u8 example [3, 255, 254] -> on u64: [3, -1, -2]
    assert list(out) == [3, U64, U64 - 1]
    # X^d then X^-d is the identity; mul_and_subtract equals rotate minus input
    rng = oracle.Rng(3)
    p = rng.uniform(64)
    for d in (0, 1, 17, 63, 64, 65, 100, 127):
        rot = np.zeros(64, dtype=np.uint64)
        L.orc_monomial_div(oracle.u64p(rot), oracle.u64p(p), 64, (128 - d) % 128)  # p * X^d
        ms = np.zeros(64, dtype=np.uint64)
        L.orc_monomial_mul_and_subtract(oracle.u64p(ms), oracle.u64p(p), 64, d)
        assert np.array_equal(ms, rot - p)
        back = np.zeros(64, dtype=np.uint64)
        L.orc_monomial_div(oracle.u64p(back), oracle.u64p(rot), 64, d)
        assert np.array_equal(back, p)


def test_modulus_switch_and_centered_correction(oracle):
    """common.rs:10-23 plus the invariant of modulus_switch.rs:55-100: the
    centered body equals body + sum(round_err/2) - half_case up to the halving
    residue (< 1 in the last bit)."""
    L = oracle.lib()
    assert L.orc_modulus_switch(C.c_uint64(0), 12) == 0
    assert L.orc_modulus_switch(C.c_uint64(U64), 12) == 0  # wraps to 0 (= 4096 mod 4096)
    assert L.orc_modulus_switch(C.c_uint64(1 << 52), 12) == 1
    assert L.orc_modulus_switch(C.c_uint64((1 << 51) - 1), 12) == 0
    assert L.orc_modulus_switch(C.c_uint64(1 << 51), 12) == 1
    rng = oracle.Rng(5)
    ct = rng.uniform(919)
    corr = L.orc_centered_ms_body_correction(oracle.u64p(ct), 918, 12)
    errs = [((int(L.orc_modulus_switch(C.c_uint64(int(a)), 12)) << 52) - int(a)) for a in ct[:918]]
    errs = [e - (1 << 64) if e >= (1 << 63) else (e + (1 << 64) if e < -(1 << 63) else e) for e in
            [((e + (1 << 63)) % (1 << 64)) - (1 << 63) for e in errs]]
    ideal = sum(errs) / 2 - (1 << 51)
    got = corr if corr < (1 << 63) else corr - (1 << 64)
    assert abs(got - ideal) <= 1.0


def test_lut_and_sample_extract(oracle):
    p = oracle.TOY_K1
    lut = oracle.make_lut(p, list(range(p.p)))
    body = lut[p.k * p.N:]
    box = p.N // p.p
    assert np.all(lut[: p.k * p.N] == 0)
    # after the half-box rotation: first half box holds f(0)*delta, last half box -f(0)*delta
    assert np.all(body[: box // 2] == 0)
    assert int(body[box // 2]) == p.delta  # f(1) * delta starts at half a box
    assert int(body[-1]) == 0  # -f(0)*delta = 0
    lut2 = oracle.make_lut(p, [3] * p.p)
    assert int(lut2[p.k * p.N + p.N - 1]) == (-3 * p.delta) % (1 << 64)
    # sample extract: decrypting the extracted LWE == coefficient nth of the GLWE phase
    rng = oracle.Rng(9)
    glwe_sk = rng.binary(p.k * p.N)
    L = oracle.lib()
    pt = rng.uniform(p.N)
    mask = np.zeros(p.k * p.N, dtype=np.uint64)
    bodyc = pt.copy()
    L.orc_glwe_encrypt_assign(rng.ref, oracle.u64p(glwe_sk), p.k, p.N, -1, oracle.u64p(mask), oracle.u64p(bodyc))
    glwe = np.concatenate([mask, bodyc])
    for nth in (0, 1, 7, p.N - 1):
        lwe = np.zeros(p.k * p.N + 1, dtype=np.uint64)
        L.orc_sample_extract(oracle.u64p(glwe), p.k, p.N, nth, oracle.u64p(lwe))
        assert int(oracle.lwe_decrypt_batch(glwe_sk, lwe[None, :])[0]) == int(pt[nth])


@pytest.mark.parametrize("pname", ["TOY_K1", "TOY_K2_L2", "TOY_MB3"])
def test_ks_pbs_semantics(oracle, keyset, pname):
    """algorithms/test/lwe_programmable_bootstrapping.rs:108-156 and
    test/lwe_keyswitch.rs: decrypt(PBS(KS(Enc m))) == f(m), FFT and exact."""
    P = getattr(oracle, pname)
    keys = keyset(P)
    rng = oracle.Rng(99)
    msgs = np.arange(2 * P.p) % P.p
    big = oracle.lwe_encrypt_batch(rng, keys.glwe_sk, msgs.astype(np.uint64) * np.uint64(P.delta), P.lwe_noise_log2)
    small = oracle.keyswitch_batch(keys, big)
    assert np.array_equal(oracle.decode(oracle.lwe_decrypt_batch(keys.lwe_sk, small), P.delta, P.p), msgs)
    f = [(3 * i + 1) % P.p for i in range(P.p)]
    lut = oracle.make_lut(P, f)
    want = np.array([f[m] for m in msgs])
    for exact in (False, True):
        out = oracle.pbs_batch(keys, lut, small, exact=exact)
        got = oracle.decode(oracle.lwe_decrypt_batch(keys.glwe_sk, out), P.delta, P.p)
        assert np.array_equal(got, want), (pname, exact)


def test_single_external_product_fft_vs_exact(oracle, keyset):
    """One CMUX (n = 1): FFT mode equals exact mode within the f64 noise floor
    (cf. pbs_golden/mod.rs:72-80: ~top 32 bits meaningful)."""
    P = oracle.Params("P22_N1", n=1, k=1, N=2048, pbs_base_log=23, pbs_level=1, ks_base_log=4, ks_level=4,
                      lwe_noise_log2=45, glwe_noise_log2=17)
    keys = keyset(P, seed=11, with_ksk=False)
    cts = oracle.lwe_encrypt_batch(oracle.Rng(3), keys.lwe_sk, (np.arange(8) % 16).astype(np.uint64) * np.uint64(P.delta), 45)
    lut = oracle.make_lut(P, list(range(16)))
    a = oracle.pbs_batch(keys, lut, cts)
    b = oracle.pbs_batch(keys, lut, cts, exact=True)
    assert np.abs((a - b).astype(np.int64)).max() < (1 << 43)


def test_many_lut_and_indexes(oracle, keyset):
    P = oracle.TOY_K1
    keys = keyset(P)
    rng = oracle.Rng(21)
    msgs = np.array([1, 2, 3, 0, 1, 2])
    # many-lut: 2 functions packed in one LUT (message space halved)
    cts = oracle.lwe_encrypt_batch(rng, keys.lwe_sk, msgs.astype(np.uint64) * np.uint64(P.delta), P.lwe_noise_log2)
    luts = np.stack([oracle.make_lut(P, [(i + 1) % P.p for i in range(P.p)]),
                     oracle.make_lut(P, [(2 * i) % P.p for i in range(P.p)])])
    lut_idx = np.array([0, 1, 0, 1, 1, 0], dtype=np.uint64)
    in_idx = np.array([5, 4, 3, 2, 1, 0], dtype=np.uint64)
    out_idx = np.array([0, 2, 4, 1, 3, 5], dtype=np.uint64)
    out = oracle.pbs_batch(keys, luts, cts, lut_idx=lut_idx, in_idx=in_idx, out_idx=out_idx)
    dec = oracle.decode(oracle.lwe_decrypt_batch(keys.glwe_sk, out), P.delta, P.p)
    for s in range(6):
        m = msgs[in_idx[s]]
        want = (m + 1) % P.p if lut_idx[s] == 0 else (2 * m) % P.p
        assert dec[out_idx[s]] == want


@pytest.mark.slow
def test_p22_full_size_semantics(oracle, keyset):
    """config[1] shape on CPU: PARAM_MESSAGE_2_CARRY_2_KS_PBS, 16 samples."""
    P = oracle.PARAM_MESSAGE_2_CARRY_2_KS_PBS
    keys = keyset(P, seed=0xB2000001)
    rng = oracle.Rng(1)
    msgs = np.arange(16)
    big = oracle.lwe_encrypt_batch(rng, keys.glwe_sk, msgs.astype(np.uint64) * np.uint64(P.delta), P.lwe_noise_log2)
    small = oracle.keyswitch_batch(keys, big)
    lut = oracle.make_lut(P, [(i * i) % 16 for i in range(16)])
    out = oracle.pbs_batch(keys, lut, small)
    got = oracle.decode(oracle.lwe_decrypt_batch(keys.glwe_sk, out), P.delta, 16)
    assert np.array_equal(got, (msgs * msgs) % 16)


BOOLEAN_SETS = {
    # tfhe/src/boolean/parameters/params.rs:10-26 (gaussian std 5.86e-6 / 9.3e-10 of the torus ~ TUniform 2^47 / 2^35)
    "DEFAULT_PARAMETERS": dict(n=805, k=3, N=512, pbs_base_log=10, pbs_level=2, ks_base_log=3, ks_level=5,
                               lwe_noise_log2=47, glwe_noise_log2=35),
    # params.rs:46-62 (the N = 1024 boolean set BASELINE.json's config[0] mentions)
    "PARAMETERS_ERROR_PROB_2_POW_MINUS_165": dict(n=837, k=2, N=1024, pbs_base_log=10, pbs_level=2, ks_base_log=3,
                                                  ks_level=5, lwe_noise_log2=46, glwe_noise_log2=35),
}


def boolean_params(oracle, name):
    return oracle.Params(name, message_bits=1, carry_bits=0, centered_ms=False, **BOOLEAN_SETS[name])


def boolean_nand_inputs(oracle, keys, rng, a, b):
    """-(ct_a + ct_b) + (0,..,0,1/8): boolean/engine/mod.rs:612-631, on the u64
    torus (the u32 torus of the boolean API embedded in the top 32 bits)."""
    P = keys.params
    eighth = 1 << 61
    enc = lambda bit: oracle.lwe_encrypt_batch(rng, keys.lwe_sk, [eighth if bit else (-eighth) % (1 << 64)],
                                               P.lwe_noise_log2)[0]
    with np.errstate(over="ignore"):
        ct = np.uint64(0) - (enc(a) + enc(b))
        ct[-1] += np.uint64(eighth)
    return ct


def boolean_lut(P):
    lut = np.zeros((P.k + 1) * P.N, dtype=np.uint64)
    lut[P.k * P.N:] = 1 << 61  # constant 1/8 accumulator (boolean/engine/bootstrapping.rs:63-64)
    return lut


@pytest.mark.parametrize("name", sorted(BOOLEAN_SETS))
def test_boolean_gate_bootstrap_on_u64_engine(oracle, keyset, name):
    """configs[0]: one boolean NAND gate bootstrap (PBS then keyswitch,
    EncryptionKeyChoice::Small, boolean/engine/bootstrapping.rs:488-532) with the
    reference's boolean parameter shapes; decrypt-equal check on CPU."""
    P = boolean_params(oracle, name)
    keys = keyset(P, seed=77)
    rng = oracle.Rng(8)
    lut = boolean_lut(P)
    for a in (0, 1):
        for b in (0, 1):
            ct = boolean_nand_inputs(oracle, keys, rng, a, b)
            big = oracle.pbs_batch(keys, lut, ct[None, :])
            small = oracle.keyswitch_batch(keys, big)
            ph = int(oracle.lwe_decrypt_batch(keys.lwe_sk, small)[0])
            bit = 1 if ph < (1 << 63) else 0
            assert bit == (0 if (a and b) else 1), (name, a, b)


@pytest.mark.slow
def test_p22_output_noise_within_reference_formula(oracle, keyset):
    """noise_distribution/lwe_programmable_bootstrapping_noise.rs:176-204: the
    measured PBS output variance must not exceed the formula by more than
    6.25 % (plus the estimator's own spread at this sample count)."""
    from tests.noise_formula import pbs_variance_tuniform_fft

    P = oracle.PARAM_MESSAGE_2_CARRY_2_KS_PBS
    keys = keyset(P, seed=0xB2000001)
    count = 192
    msgs = np.arange(count) % 16
    cts = oracle.lwe_encrypt_batch(oracle.Rng(1), keys.lwe_sk, msgs.astype(np.uint64) * np.uint64(P.delta),
                                   P.lwe_noise_log2)
    out = oracle.pbs_batch(keys, oracle.make_lut(P, list(range(16))), cts)
    ph = oracle.lwe_decrypt_batch(keys.glwe_sk, out)
    noise = (ph - msgs.astype(np.uint64) * np.uint64(P.delta)).astype(np.int64).astype(np.float64) / 2.0 ** 64
    bound = pbs_variance_tuniform_fft(P.n, P.k, P.N, P.pbs_base_log, P.pbs_level)
    spread = 1.0 + 4.0 * np.sqrt(2.0 / (count - 1))
    assert noise.var() < bound * 1.0625 * spread, (noise.var(), bound)
    assert noise.var() > bound * 0.5  # sanity: same order as the prediction
