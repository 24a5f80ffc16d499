"""CPU: known-answer tests of oracle/glm_shim -- the stand-in for GLM (a vcpkg dependency that is not vendored
under /root/reference) against which the reference's kernels are compiled in oracle/_ref.  Every
b200-vs-reference parity test inherits the shim's semantics, and the oracle restates the same conventions, so the
shim is pinned here against implementations this repository did not write (scipy.spatial.transform, numpy.linalg)
and against closed-form answers: column-major storage, (w, x, y, z) quaternions, q * v, quat_cast / mat3_cast,
slerp (endpoints, midpoint, shortest arc, identical endpoints), inverse(mat2), outerProduct."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from scipy.spatial.transform import Rotation, Slerp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "glm_kat", "glm_kat.cpp")
OUT = os.path.join(ROOT, "tests", "glm_kat", "libglm_kat.so")


@pytest.fixture(scope="module")
def kat():
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "oracle", "glm_shim"),
                        SRC, "-o", OUT], check=True)
    L = C.CDLL(OUT)
    L.kat_dot3.restype = C.c_float
    L.kat_length3.restype = C.c_float
    return L


def f(a):
    return np.ascontiguousarray(a, np.float32)


def p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def call(fn, *ins, n):
    out = np.zeros(n, np.float32)
    keep = [x if np.isscalar(x) else f(x) for x in ins]  # the arrays must outlive the call
    fn(*[C.c_float(x) if np.isscalar(x) else p(x) for x in keep], p(out))
    return out


def colmajor(M):  # math matrix -> GLM storage order
    return f(np.asarray(M).T.reshape(-1))


def from_colmajor(v, n=3):
    return np.asarray(v, np.float64).reshape(n, n).T


def scipy_rot(q_wxyz):
    return Rotation.from_quat([q_wxyz[1], q_wxyz[2], q_wxyz[3], q_wxyz[0]])


def test_mat3_cast_and_rotate_closed_form(kat):
    s = np.sqrt(0.5)
    qz = [s, 0, 0, s]  # 90 degrees about +z
    R = from_colmajor(call(kat.kat_mat3_cast, qz, n=9))
    np.testing.assert_allclose(R, [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-6)
    np.testing.assert_allclose(call(kat.kat_rotate, qz, [1, 0, 0], n=3), [0, 1, 0], atol=1e-6)
    np.testing.assert_allclose(call(kat.kat_quat_mul_vec, [s, s, 0, 0], [0, 1, 0], n=3), [0, 0, 1], atol=1e-6)  # about +x
    np.testing.assert_allclose(call(kat.kat_quat_mul_vec, [s, 0, s, 0], [0, 0, 1], n=3), [1, 0, 0], atol=1e-6)  # about +y
    # storage: element [col][row]; the first three floats are the first COLUMN
    M = np.arange(9, dtype=np.float32).reshape(3, 3) + 1
    np.testing.assert_allclose(call(kat.kat_mat3_mul_vec, colmajor(M), [1, 0, 0], n=3), M[:, 0])
    np.testing.assert_allclose(from_colmajor(call(kat.kat_transpose3, colmajor(M), n=9)), M.T)


def test_quaternion_functions_vs_scipy(kat):
    rng = np.random.default_rng(0)
    for _ in range(200):
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        v = rng.standard_normal(3)
        R = from_colmajor(call(kat.kat_mat3_cast, q, n=9))
        np.testing.assert_allclose(R, scipy_rot(q).as_matrix(), atol=2e-6)
        np.testing.assert_allclose(call(kat.kat_rotate, q, v, n=3), scipy_rot(q).apply(v), atol=5e-6)
        # quat_cast inverts mat3_cast up to the sign of the quaternion, for every branch of the "biggest" switch
        q2 = call(kat.kat_quat_cast, colmajor(scipy_rot(q).as_matrix()), n=4)
        assert min(np.abs(q2 - q).max(), np.abs(q2 + q).max()) < 3e-6
        # inverse of a non-unit quaternion is conjugate / |q|^2
        qs = q * 1.7
        qi = call(kat.kat_inverse_quat, qs, n=4)
        np.testing.assert_allclose(qi, np.array([qs[0], -qs[1], -qs[2], -qs[3]]) / np.dot(qs, qs), atol=1e-6)
        np.testing.assert_allclose(call(kat.kat_normalize_quat, qs, n=4), q, atol=1e-6)
    # all four branches of quat_cast (trace-dominant, x-, y-, z-dominant)
    for axis, ang in (((1, 0, 0), 3.0), ((0, 1, 0), 3.0), ((0, 0, 1), 3.0), ((1, 1, 1), 0.3)):
        r = Rotation.from_rotvec(np.array(axis, float) / np.linalg.norm(axis) * ang)
        q = call(kat.kat_quat_cast, colmajor(r.as_matrix()), n=4)
        x, y, z, w = r.as_quat()
        want = np.array([w, x, y, z])
        assert min(np.abs(q - want).max(), np.abs(q + want).max()) < 3e-6


def test_slerp(kat):
    rng = np.random.default_rng(1)
    for _ in range(100):
        a, b = rng.standard_normal(4), rng.standard_normal(4)
        a /= np.linalg.norm(a)
        b /= np.linalg.norm(b)
        np.testing.assert_allclose(call(kat.kat_slerp, a, b, 0.0, n=4), a, atol=2e-6)
        end = call(kat.kat_slerp, a, b, 1.0, n=4)
        assert min(np.abs(end - b).max(), np.abs(end + b).max()) < 2e-6  # shortest arc may flip the sign of b
        t = float(rng.random())
        got = call(kat.kat_slerp, a, b, t, n=4)
        want = Slerp([0, 1], Rotation.concatenate([scipy_rot(a), scipy_rot(b)]))([t])
        np.testing.assert_allclose(from_colmajor(call(kat.kat_mat3_cast, got, n=9)), want.as_matrix()[0], atol=1e-5)
        # identical endpoints return the start quaternion for any t: what the reference's global shutter relies on
        # (Cameras.cuh:268-280)
        np.testing.assert_allclose(call(kat.kat_slerp, a, a, t, n=4), a, atol=1e-6)


def test_small_matrix_functions(kat):
    # inverse(mat2): [[4, 7], [2, 6]]^-1 = [[0.6, -0.7], [-0.2, 0.4]]
    M = np.array([[4.0, 7.0], [2.0, 6.0]])
    inv = from_colmajor(call(kat.kat_inverse_mat2, f(M.T.reshape(-1)), n=4), 2)
    np.testing.assert_allclose(inv, [[0.6, -0.7], [-0.2, 0.4]], atol=1e-6)
    rng = np.random.default_rng(2)
    for _ in range(50):
        M = rng.standard_normal((2, 2)) + 2 * np.eye(2)
        inv = from_colmajor(call(kat.kat_inverse_mat2, f(M.T.reshape(-1)), n=4), 2)
        np.testing.assert_allclose(inv, np.linalg.inv(M), rtol=2e-5, atol=2e-6)
        A, B = rng.standard_normal((3, 3)), rng.standard_normal((3, 3))
        c, r = rng.standard_normal(3), rng.standard_normal(3)
        np.testing.assert_allclose(from_colmajor(call(kat.kat_mat3_mul_mat3, colmajor(A), colmajor(B), n=9)), A @ B,
                                   atol=1e-5)
        np.testing.assert_allclose(call(kat.kat_mat3_mul_vec, colmajor(A), c, n=3), A @ c, atol=1e-5)
        np.testing.assert_allclose(from_colmajor(call(kat.kat_outer3, c, r, n=9)), np.outer(c, r), atol=1e-6)
        np.testing.assert_allclose(call(kat.kat_cross, c, r, n=3), np.cross(c, r), atol=1e-6)
        np.testing.assert_allclose(call(kat.kat_normalize3, c, n=3), c / np.linalg.norm(c), atol=1e-6)
        assert abs(kat.kat_dot3(p(f(c)), p(f(r))) - float(np.dot(c, r))) < 1e-5
        assert abs(kat.kat_length3(p(f(c))) - float(np.linalg.norm(c))) < 1e-5
