"""Oracle self-checks: the scalar "virtual lane chain" emulations of the reference's explicit SIMD kernels are
bit-identical to the same kernels written with the reference's intrinsics (src/arch/x86_64.rs:31-106, 183-265,
681-786, 799-915, 928-1020), at every tail shape; generators reproduce the values recorded in SURVEY.md F11;
backend strings (src/backend.rs:96-120)."""
import numpy as np
import pytest

DIMS = [1, 7, 8, 15, 16, 17, 31, 32, 33, 63, 64, 65, 79, 80, 96, 127, 128, 129, 200, 384, 767, 768, 1535, 1536]


def _rand(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n) * rng.choice([1e-3, 1.0, 50.0])).astype(np.float32)


@pytest.mark.parametrize("name", ["dot_avx512", "dot_avx2", "cosine_avx512", "cosine_avx2"])
def test_emulation_is_bit_identical_to_intrinsics(oracle, name):
    if not oracle.host_has_avx512():
        pytest.skip("host lacks AVX-512F: only the emulation can run here")
    intr, emul = getattr(oracle, name + "_intrin"), getattr(oracle, name + "_emul")
    for dim in DIMS:
        for seed in range(6):
            a, b = _rand(dim, seed), _rand(dim, seed + 100)
            x, y = np.float32(intr(a, b)), np.float32(emul(a, b))
            assert x.tobytes() == y.tobytes(), (name, dim, seed, x, y)


def test_u8_emulation_is_bit_identical_to_intrinsics(oracle):
    rng = np.random.default_rng(7)
    for dim in DIMS:
        for seed in range(6):
            a = _rand(dim, seed)
            b = rng.integers(0, 256, dim, dtype=np.uint8)
            x = np.float32(oracle.dot_u8_f32_variant("dot_u8_f32_avx2_intrin", a, b))
            y = np.float32(oracle.dot_u8_f32_variant("dot_u8_f32_avx2_emul", a, b))
            assert x.tobytes() == y.tobytes(), (dim, seed)


def test_forced_emulation_mode_changes_nothing(oracle):
    q = np.stack([_rand(128, i) for i in range(4)])
    d = np.stack([_rand(128, 50 + i) for i in range(9)])
    a = oracle.maxsim(q, d), oracle.maxsim_cosine(q, d)
    oracle.set_simd_mode(1)
    try:
        b = oracle.maxsim(q, d), oracle.maxsim_cosine(q, d)
    finally:
        oracle.set_simd_mode(0)
    assert np.float32(a[0]).tobytes() == np.float32(b[0]).tobytes()
    assert np.float32(a[1]).tobytes() == np.float32(b[1]).tobytes()


def test_dispatch_thresholds(oracle):  # src/dense.rs:70-100: n>=64 avx512, n>=16 avx2, else portable
    for n, fn in ((8, oracle.dot_portable), (15, oracle.dot_portable), (16, oracle.dot_avx2_emul),
                  (63, oracle.dot_avx2_emul), (64, oracle.dot_avx512_emul), (768, oracle.dot_avx512_emul)):
        a, b = _rand(n, n), _rand(n, n + 1)
        assert np.float32(oracle.dot(a, b)).tobytes() == np.float32(fn(a, b)).tobytes()


def test_batch_dot_is_sequential_unfused(oracle):  # strict sequential f32 sum, mul and add rounded apart
    n, d = 37, 768
    rows = np.stack([_rand(d, i) for i in range(n)])
    q = _rand(d, 999)
    b = oracle.VerticalBatch.from_flat(rows.reshape(-1), n, d)
    got = oracle.batch_dot(q, b)
    for i in range(n):
        acc = np.float32(0)
        for dd in range(d):
            acc = np.float32(acc + np.float32(q[dd] * rows[i, dd]))
        assert acc.tobytes() == got[i].tobytes()


def test_generate_embedding_known_values(oracle):  # v(0,.) = -1.0, -0.8436, -0.6872, ...
    v = oracle.generate_embedding(8, 0)
    assert v[0] == -1.0 and abs(v[1] + 0.8436) < 1e-4 and abs(v[2] + 0.6872) < 1e-4
    # pure-Python restatement of examples/batch_demo.rs:233-242
    for seed in (0, 1, 999, 50_000, 2**40 + 3):
        got = oracle.generate_embedding(16, seed)
        for i in range(16):
            x = (seed * 6364136223846793005 + i * 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
            want = np.float32(np.float32(np.float32(x >> 33) / np.float32(2**31)) * np.float32(2.0) - np.float32(1.0))
            assert got[i].tobytes() == want.tobytes()
    n = oracle.generate_normalized(128, 5)
    assert abs(float(np.sum(n.astype(np.float64) ** 2)) - 1.0) < 1e-5


def test_ghash_generator(oracle):
    v = oracle.ghash_f32(0x5EED0000, 0, 4096)
    assert v.min() >= -1.0 and v.max() < 1.0 and abs(float(v.mean())) < 0.05
    u = oracle.ghash_u64(0x5EED0002, 10, 3)
    assert int(u[0]) == oracle.splitmix64(0x5EED0002 + 10)
    # splitmix64 known answer (reference implementation by Vigna, seed 0 first output)
    assert oracle.splitmix64(0) == 0xE220A8397B1DCDAF


def test_backend_strings(oracle):  # src/backend.rs:96-120
    assert oracle.dense_backend(1) == "portable" and oracle.dense_backend(15) == "portable"
    assert oracle.dense_backend(768) in ("avx512", "avx2+fma", "portable")
    if oracle.host_has_avx512():
        assert oracle.dense_backend(768) == "avx512" and oracle.dense_backend(32) == "avx2+fma"
