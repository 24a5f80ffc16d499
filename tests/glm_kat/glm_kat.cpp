// Known-answer harness for oracle/glm_shim (test infrastructure).  The reference's device code needs GLM
// (vcpkg dependency, not vendored under /root/reference); oracle/glm_shim supplies the functions the reference
// calls, and every b200-vs-reference test therefore depends on it.  This file exposes each of those functions
// through a C ABI so that tests/test_glm_shim_kat.py can check them against independent implementations
// (scipy.spatial.transform, numpy) and closed-form answers.  Matrices cross the ABI column-major (GLM's layout).
#include <glm/glm.hpp>
#include <glm/gtc/quaternion.hpp>
#include <glm/gtc/type_ptr.hpp>
#include <glm/gtx/matrix_operation.hpp>
#include <glm/gtx/quaternion.hpp>

static glm::fquat q_in(const float *q) { return glm::fquat(q[0], q[1], q[2], q[3]); } // (w, x, y, z)
static void q_out(const glm::fquat &q, float *o) { o[0] = q.w; o[1] = q.x; o[2] = q.y; o[3] = q.z; }
static glm::fmat3 m3_in(const float *m) { return glm::fmat3(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8]); }
static void m3_out(const glm::fmat3 &m, float *o) {
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) o[c * 3 + r] = m[c][r];
}

extern "C" {
void kat_mat3_cast(const float *q, float *m) { m3_out(glm::mat3_cast(q_in(q)), m); }
void kat_quat_cast(const float *m, float *q) { q_out(glm::quat_cast(m3_in(m)), q); }
void kat_rotate(const float *q, const float *v, float *o) {
    const glm::fvec3 r = glm::rotate(q_in(q), glm::make_vec3(v));
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
void kat_quat_mul_vec(const float *q, const float *v, float *o) {
    const glm::fvec3 r = q_in(q) * glm::make_vec3(v);
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
void kat_inverse_quat(const float *q, float *o) { q_out(glm::inverse(q_in(q)), o); }
void kat_normalize_quat(const float *q, float *o) { q_out(glm::normalize(q_in(q)), o); }
void kat_slerp(const float *a, const float *b, float t, float *o) { q_out(glm::slerp(q_in(a), q_in(b), t), o); }
void kat_inverse_mat2(const float *m, float *o) {
    const glm::fmat2 r = glm::inverse(glm::fmat2(m[0], m[1], m[2], m[3]));
    o[0] = r[0][0]; o[1] = r[0][1]; o[2] = r[1][0]; o[3] = r[1][1];
}
void kat_mat3_mul_vec(const float *m, const float *v, float *o) {
    const glm::fvec3 r = m3_in(m) * glm::make_vec3(v);
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
void kat_mat3_mul_mat3(const float *a, const float *b, float *o) { m3_out(m3_in(a) * m3_in(b), o); }
void kat_transpose3(const float *m, float *o) { m3_out(glm::transpose(m3_in(m)), o); }
void kat_outer3(const float *c, const float *r, float *o) { m3_out(glm::outerProduct(glm::make_vec3(c), glm::make_vec3(r)), o); }
void kat_cross(const float *a, const float *b, float *o) {
    const glm::fvec3 r = glm::cross(glm::make_vec3(a), glm::make_vec3(b));
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
float kat_dot3(const float *a, const float *b) { return glm::dot(glm::make_vec3(a), glm::make_vec3(b)); }
void kat_normalize3(const float *a, float *o) {
    const glm::fvec3 r = glm::normalize(glm::make_vec3(a));
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
float kat_length3(const float *a) { return glm::length(glm::make_vec3(a)); }
}
