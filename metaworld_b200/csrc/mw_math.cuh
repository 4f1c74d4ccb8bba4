// Small vector / quaternion helpers for the step kernel (row-major 3x3, quaternion w,x,y,z).
#pragma once
#ifndef MW_HOST_EMU   // tests/devcollide compiles the collision code for the host with one emulated lane
#include <cuda_runtime.h>
#endif

#ifndef MW_REAL_DOUBLE
typedef float real;
#define R_(x) x##f
#define MW_EPS 1e-12f
#else
typedef double real;
#define R_(x) x
#define MW_EPS 1e-15
#endif

#define DEV __device__ __forceinline__
#define FULLMASK 0xffffffffu

// creal: scalar type of the collision geometry (always float64, like the reference's mjtNum: contact existence is a
// discontinuity of the dynamics, so the narrowphase must not be the place where float32 rounding decides)
typedef double creal;
template <class T> struct ty_ { typedef T type; };
template <class T> DEV T eps_() { return (T)1e-15; }
template <> DEV float eps_<float>() { return 1e-12f; }

DEV real rsqrt_(real x) { return (real)1 / sqrt(x); }
template <class T> DEV void v3zero(T* a) { a[0] = a[1] = a[2] = 0; }
template <class T> DEV void v3copy(T* a, const T* b) { a[0] = b[0]; a[1] = b[1]; a[2] = b[2]; }
template <class T> DEV void v3add(T* r, const T* a, const T* b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
template <class T> DEV void v3sub(T* r, const T* a, const T* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
template <class T> DEV void v3scl(T* r, const T* a, typename ty_<T>::type s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
template <class T> DEV void v3addscl(T* r, const T* a, const T* b, typename ty_<T>::type s) { r[0] = a[0] + b[0] * s; r[1] = a[1] + b[1] * s; r[2] = a[2] + b[2] * s; }
template <class T> DEV T v3dot(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class T> DEV T v3norm(const T* a) { return sqrt(v3dot(a, a)); }
template <class T> DEV void v3cross(T* r, const T* a, const T* b) {
  T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
template <class T> DEV T v3normalize(T* a) {
  T n = v3norm(a);
  if (n < eps_<T>()) { a[0] = 1; a[1] = a[2] = 0; return 0; }
  T s = (T)1 / n; a[0] *= s; a[1] *= s; a[2] *= s; return n;
}
template <class T> DEV void mat_mulvec(T* r, const T* M, const T* v) {
  T x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2], y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2], z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <class T> DEV void mat_tmulvec(T* r, const T* M, const T* v) {
  T x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2], y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2], z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <class T> DEV void mat_col(T* r, const T* M, int c) { r[0] = M[c]; r[1] = M[3 + c]; r[2] = M[6 + c]; }
template <class T> DEV void mat_mul(T* r, const T* A, const T* B) {
  T t[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
#pragma unroll
  for (int i = 0; i < 9; i++) r[i] = t[i];
}
template <class T> DEV void quat_mul(T* r, const T* a, const T* b) {
  T w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  T x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  T y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  T z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
template <class T> DEV void quat_normalize(T* q) {
  T n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < eps_<T>()) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  T s = (T)1 / n; q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
}
template <class T> DEV void quat2mat(T* R, const T* q) {
  T w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}
template <class T> DEV void quat_rot(T* r, const T* q, const T* v) { T R[9]; quat2mat(R, q); mat_mulvec(r, R, v); }
template <class T> DEV void quat_axisangle(T* q, const T* axis, typename ty_<T>::type ang) {
  T s, c; sincos((T)0.5 * ang, &s, &c);
  q[0] = c; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
template <class T> DEV T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}
template <class T> DEV T bcast(T v, int src) { return __shfl_sync(FULLMASK, v, src); }
