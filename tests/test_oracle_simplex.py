"""The C simplex oracle vs golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  Bit-exact: integer tables and fp64 values."""
import hashlib

import numpy as np

from oracle.simplex_oracle import OracleSimplex, init_tables


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_perm_tables_bit_exact(golden):
    g = golden("simplex")
    for i, s in enumerate(g["seeds"]):
        perm, pgi3 = init_tables(int(s))
        assert np.array_equal(perm, g["perm"][i].astype(np.int64)), s
        assert np.array_equal(pgi3, g["pgi3"][i].astype(np.int64)), s


def test_survey_known_answers():
    # SURVEY.md Appendix C.3 (values produced by the reference at survey time)
    p, q = init_tables(3)
    assert list(p[:6]) == [164, 187, 231, 144, 104, 73] and list(q[:6]) == [60, 57, 45, 0, 24, 3]
    assert hashlib.sha1(p.tobytes()).hexdigest()[:16] == "4ea0ad7f9c1b383f"
    S = OracleSimplex(12345)
    assert S.noise3(0.5, 0.25, 0.125) == -0.05566569507584032
    assert S.noise3(1, 1, 1) == 0.09061488673139152
    assert S.noise3(100.3, -4.2, 55.5) == -0.5339494132559855
    assert S.noise3(0, 0, 0) == -1.4443096571220408e-66


def test_point_values_bit_exact(golden):
    g = golden("simplex")
    S = OracleSimplex(12345)
    P = g["points"]
    n3 = np.array([S.noise3(*p) for p in P])
    n2 = np.array([S.noise2(p[0], p[1]) for p in P])
    assert np.array_equal(n3, g["noise3"])
    assert np.array_equal(n2, g["noise2"])


def test_octave_fields_bit_exact(golden):
    g = golden("simplex")
    S = OracleSimplex(12345, threads=4)
    assert np.array_equal(S.rand_3d_fixed_T_octaves((64, 64), np.array([17]), 6, 0.8, 64), g["fixedT_64_t17"])
    assert np.array_equal(S.rand_3d_fixed_T_octaves((48, 80), np.array([999]), 3, 0.6, 16), g["fixedT_48x80_t999_o3"])
    big = S.rand_3d_fixed_T_octaves((256, 256), np.array([500]), 6, 0.8, 64)
    assert sha(big) == str(g["fixedT_256_t500_sha1"])
    assert sha(big.astype(np.float32)) == str(g["fixedT_256_t500_f32_sha1"])
    assert abs(big.sum() - 2332.670612811888) < 1e-9
    assert np.array_equal(S.rand_3d_octaves((4, 32, 32), 8, 0.8, 64), g["vol_4x32x32_o8"])
    assert np.array_equal(S.noise3array(np.arange(9) / 3.0, np.arange(7) / 5.0, np.arange(5) / 7.0),
                          g["noise3array_5x7x9"])


def test_config5_volume_bit_exact(golden):
    g = golden("simplex")
    S = OracleSimplex(12345, threads=8)
    vol = S.rand_3d_octaves((32, 256, 256), 8, 0.8, 64)
    assert sha(vol) == str(g["vol_32x256x256_o8_sha1"])
    assert np.array_equal(vol[::8, 120:136, 64:96], g["vol_32x256x256_o8_crop"])


def test_host_seed_stream(golden):
    g = golden("simplex")
    np.random.seed(0)
    draws = [np.random.randint(-10000000000, 10000000000) for _ in range(8)]
    assert draws == list(g["np_seed0_draws"])
    assert draws[:3] == [9726117423, 6569750267, 6523820391]
