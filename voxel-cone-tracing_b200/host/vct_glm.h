// vct_glm.h -- the handful of glm functions the reference's host code uses (glm itself is not vendored in the
// reference and is absent from this image): vec3, mat4 (column-major, m[col][row] like glm), lookAt, ortho,
// perspective, scale, radians.  Right-handed, NDC z in [-1, 1] -- glm's defaults.
#pragma once
#include <cmath>

namespace vctm {

struct vec3 {
  float x, y, z;
  vec3() : x(0), y(0), z(0) {}
  vec3(float a, float b, float c) : x(a), y(b), z(c) {}
  explicit vec3(float a) : x(a), y(a), z(a) {}
};
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline vec3 normalize(vec3 a) { float l = std::sqrt(dot(a, a)); return {a.x / l, a.y / l, a.z / l}; }

struct mat4 {
  float m[4][4];   // m[column][row]
  mat4() : mat4(1.0f) {}
  explicit mat4(float d) { for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) m[c][r] = (c == r) ? d : 0.0f; }
  const float* data() const { return &m[0][0]; }   // 16 floats, column-major: what glUniformMatrix4fv(.., GL_FALSE, ..) takes
};
inline mat4 operator*(const mat4& a, const mat4& b) {
  mat4 r(0.0f);
  for (int c = 0; c < 4; ++c)
    for (int rr = 0; rr < 4; ++rr) {
      float s = 0.0f;
      for (int k = 0; k < 4; ++k) s += a.m[k][rr] * b.m[c][k];
      r.m[c][rr] = s;
    }
  return r;
}
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
inline mat4 scale(const mat4& m, vec3 s) {
  mat4 r = m;
  for (int rr = 0; rr < 4; ++rr) { r.m[0][rr] *= s.x; r.m[1][rr] *= s.y; r.m[2][rr] *= s.z; }
  return r;
}
inline mat4 translate(const mat4& m, vec3 t) {
  mat4 r = m;
  for (int rr = 0; rr < 4; ++rr) r.m[3][rr] = m.m[0][rr] * t.x + m.m[1][rr] * t.y + m.m[2][rr] * t.z + m.m[3][rr];
  return r;
}
inline mat4 lookAt(vec3 eye, vec3 center, vec3 up) {
  vec3 f = normalize(center - eye), s = normalize(cross(f, up)), u = cross(s, f);
  mat4 r(1.0f);
  r.m[0][0] = s.x; r.m[1][0] = s.y; r.m[2][0] = s.z;
  r.m[0][1] = u.x; r.m[1][1] = u.y; r.m[2][1] = u.z;
  r.m[0][2] = -f.x; r.m[1][2] = -f.y; r.m[2][2] = -f.z;
  r.m[3][0] = -dot(s, eye); r.m[3][1] = -dot(u, eye); r.m[3][2] = dot(f, eye);
  return r;
}
inline mat4 ortho(float l, float r_, float b, float t, float n, float f) {
  mat4 r(1.0f);
  r.m[0][0] = 2.0f / (r_ - l); r.m[1][1] = 2.0f / (t - b); r.m[2][2] = -2.0f / (f - n);
  r.m[3][0] = -(r_ + l) / (r_ - l); r.m[3][1] = -(t + b) / (t - b); r.m[3][2] = -(f + n) / (f - n);
  return r;
}
inline mat4 perspective(float fovy, float aspect, float n, float f) {
  float th = std::tan(fovy / 2.0f);
  mat4 r(0.0f);
  r.m[0][0] = 1.0f / (aspect * th); r.m[1][1] = 1.0f / th; r.m[2][2] = -(f + n) / (f - n);
  r.m[2][3] = -1.0f; r.m[3][2] = -(2.0f * f * n) / (f - n);
  return r;
}

}  // namespace vctm
