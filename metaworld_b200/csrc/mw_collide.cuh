// Narrowphase for the step kernel.
//
// Conventions follow the reference physics (MuJoCo mj_collision [3P], reached from
// metaworld/sawyer_xyz_env.py:595,620): geom1 has the lower type id, the normal points from
// geom1 to geom2, dist < 0 is penetration, pos is the midpoint between the two surfaces.
//
// * analytic pairs (plane-*, sphere/capsule pairs, sphere-box, capsule-box, box-box) are
//   evaluated ONE PAIR PER LANE;
// * general convex pairs (cylinder-*, mesh-*) run GJK + EPA with lane 0 as the leader that owns
//   the simplex / polytope (kept in the warp's shared-memory workspace) while every support
//   query is evaluated by the whole warp (mesh hull vertices are split across lanes).
#pragma once
#include "mw_math.cuh"

enum { G_PLANE = 0, G_HFIELD, G_SPHERE, G_CAPSULE, G_ELLIPSOID, G_CYLINDER, G_BOX, G_MESH };

struct RawCon { real dist, pos[3], normal[3]; };

struct DShape {
  int type; real pos[3]; real mat[9]; real size[3]; const float* vert; int nvert;
};

// ------------------------------------------------------------------ plane pairs
DEV int plane_sphere_raw(const real* pn, const real* pp, const real* c, real r, real margin, RawCon* o) {
  real t[3]; v3sub(t, c, pp);
  real dist = v3dot(t, pn) - r;
  if (dist > margin) return 0;
  o->dist = dist; v3copy(o->normal, pn);
  v3addscl(o->pos, c, pn, -(r + (real)0.5 * dist));
  return 1;
}
DEV int plane_capsule(const DShape& a, const DShape& b, real margin, RawCon* o) {
  real n[3], ax[3], c[3]; mat_col(n, a.mat, 2); mat_col(ax, b.mat, 2);
  int cnt = 0;
  v3addscl(c, b.pos, ax, b.size[1]); cnt += plane_sphere_raw(n, a.pos, c, b.size[0], margin, o + cnt);
  v3addscl(c, b.pos, ax, -b.size[1]); cnt += plane_sphere_raw(n, a.pos, c, b.size[0], margin, o + cnt);
  return cnt;
}
DEV int plane_cylinder(const DShape& a, const DShape& b, real margin, RawCon* o) {
  real n[3], axis[3], vec[3], t[3];
  mat_col(n, a.mat, 2); mat_col(axis, b.mat, 2);
  v3sub(t, b.pos, a.pos);
  real dist0 = v3dot(t, n), prjaxis = v3dot(n, axis);
  if (prjaxis > 0) { v3scl(axis, axis, -1); prjaxis = -prjaxis; }
  for (int i = 0; i < 3; i++) vec[i] = prjaxis * axis[i] - n[i];
  real len = v3norm(vec);
  if (len >= (real)1e-6) v3scl(vec, vec, b.size[0] / len);
  else { mat_col(vec, b.mat, 0); v3scl(vec, vec, b.size[0]); }
  real prjvec = v3dot(vec, n);
  v3scl(axis, axis, b.size[1]); prjaxis *= b.size[1];
  int cnt = 0;
  if (dist0 + prjaxis + prjvec > margin) return 0;
  real dd = dist0 + prjaxis + prjvec;
  o[cnt].dist = dd; v3copy(o[cnt].normal, n);
  for (int i = 0; i < 3; i++) o[cnt].pos[i] = b.pos[i] + vec[i] + axis[i] - n[i] * dd * (real)0.5;
  cnt++;
  if (dist0 - prjaxis + prjvec <= margin) {
    dd = dist0 - prjaxis + prjvec;
    o[cnt].dist = dd; v3copy(o[cnt].normal, n);
    for (int i = 0; i < 3; i++) o[cnt].pos[i] = b.pos[i] + vec[i] - axis[i] - n[i] * dd * (real)0.5;
    cnt++;
  }
  real prjvec1 = -prjvec * (real)0.5;
  if (dist0 + prjaxis + prjvec1 <= margin) {
    real vec1[3]; v3cross(vec1, vec, axis); v3normalize(vec1); v3scl(vec1, vec1, b.size[0] * (real)0.8660254037844386);
    dd = dist0 + prjaxis + prjvec1;
    for (int s = -1; s <= 1; s += 2) {
      o[cnt].dist = dd; v3copy(o[cnt].normal, n);
      for (int i = 0; i < 3; i++) o[cnt].pos[i] = b.pos[i] + s * vec1[i] + axis[i] - vec[i] * (real)0.5 - n[i] * dd * (real)0.5;
      cnt++;
    }
  }
  return cnt;
}
DEV int plane_box(const DShape& a, const DShape& b, real margin, RawCon* o) {
  real n[3], t[3]; mat_col(n, a.mat, 2); v3sub(t, b.pos, a.pos);
  real dist = v3dot(t, n);
  int cnt = 0;
  for (int i = 0; i < 8 && cnt < 4; i++) {
    real v[3] = {(i & 1 ? 1 : -1) * b.size[0], (i & 2 ? 1 : -1) * b.size[1], (i & 4 ? 1 : -1) * b.size[2]}, c[3];
    mat_mulvec(c, b.mat, v);
    real ld = v3dot(n, c);
    if (dist + ld > margin) continue;
    o[cnt].dist = dist + ld; v3copy(o[cnt].normal, n);
    for (int k = 0; k < 3; k++) o[cnt].pos[k] = b.pos[k] + c[k] - n[k] * o[cnt].dist * (real)0.5;
    cnt++;
  }
  return cnt;
}

// ------------------------------------------------------------------ sphere / capsule pairs
DEV int sphere_sphere_raw(const real* c1, real r1, const real* c2, real r2, real margin, RawCon* o) {
  real dif[3]; v3sub(dif, c2, c1);
  real len = v3norm(dif), dist = len - r1 - r2;
  if (dist > margin) return 0;
  if (len < MW_EPS) { dif[0] = 1; dif[1] = dif[2] = 0; } else v3scl(dif, dif, 1 / len);
  o->dist = dist; v3copy(o->normal, dif);
  v3addscl(o->pos, c1, dif, r1 + (real)0.5 * dist);
  return 1;
}
DEV int sphere_capsule(const DShape& a, const DShape& b, real margin, RawCon* o) {
  real ax[3], t[3], c[3]; mat_col(ax, b.mat, 2); v3sub(t, a.pos, b.pos);
  real x = v3dot(t, ax);
  x = fmin(fmax(x, -b.size[1]), b.size[1]);
  v3addscl(c, b.pos, ax, x);
  return sphere_sphere_raw(a.pos, a.size[0], c, b.size[0], margin, o);
}
DEV int capsule_capsule(const DShape& a, const DShape& b, real margin, RawCon* o) {
  real a1[3], a2[3], dif[3]; mat_col(a1, a.mat, 2); mat_col(a2, b.mat, 2); v3sub(dif, a.pos, b.pos);
  real h1 = a.size[1], h2 = b.size[1];
  real ma = v3dot(a1, a1), mb = -v3dot(a1, a2), mc = v3dot(a2, a2), u = -v3dot(a1, dif), v = v3dot(a2, dif);
  real det = ma * mc - mb * mb;
  if (fabs(det) >= (real)1e-6) {
    real x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
    if (x1 > h1) { x1 = h1; x2 = (v - mb * h1) / mc; } else if (x1 < -h1) { x1 = -h1; x2 = (v + mb * h1) / mc; }
    if (x2 > h2) { x2 = h2; x1 = fmin(fmax((u - mb * h2) / ma, -h1), h1); }
    else if (x2 < -h2) { x2 = -h2; x1 = fmin(fmax((u + mb * h2) / ma, -h1), h1); }
    real p1[3], p2[3]; v3addscl(p1, a.pos, a1, x1); v3addscl(p2, b.pos, a2, x2);
    return sphere_sphere_raw(p1, a.size[0], p2, b.size[0], margin, o);
  }
  int cnt = 0;
  for (int k = 0; k < 2; k++) {
    real p1[3], p2[3], t[3]; v3addscl(p1, a.pos, a1, k ? -h1 : h1); v3sub(t, p1, b.pos);
    real x2 = fmin(fmax(v3dot(t, a2), -h2), h2);
    v3addscl(p2, b.pos, a2, x2);
    cnt += sphere_sphere_raw(p1, a.size[0], p2, b.size[0], margin, o + cnt);
  }
  return cnt;
}
DEV int sphere_box_raw(const real* c, real r, const DShape& b, real margin, RawCon* o) {
  real t[3], cl[3], clamped[3]; v3sub(t, c, b.pos); mat_tmulvec(cl, b.mat, t);
  bool inside = true;
  for (int i = 0; i < 3; i++) {
    clamped[i] = cl[i];
    if (clamped[i] > b.size[i]) { clamped[i] = b.size[i]; inside = false; }
    else if (clamped[i] < -b.size[i]) { clamped[i] = -b.size[i]; inside = false; }
  }
  real nl[3], dist, pl[3];
  if (!inside) {
    v3sub(nl, clamped, cl);
    real len = v3normalize(nl);
    dist = len - r;
    if (dist > margin) return 0;
    v3addscl(pl, cl, nl, r + (real)0.5 * dist);
  } else {
    int k = 0; real best = (real)1e30, sgn = 1;
    for (int i = 0; i < 3; i++) {
      real dpos = b.size[i] - cl[i], dneg = b.size[i] + cl[i];
      if (dpos < best) { best = dpos; k = i; sgn = 1; }
      if (dneg < best) { best = dneg; k = i; sgn = -1; }
    }
    v3zero(nl); nl[k] = -sgn;
    dist = -(best + r);
    v3copy(pl, cl); pl[k] = cl[k] + sgn * (real)0.5 * (best - r);
  }
  o->dist = dist;
  mat_mulvec(o->normal, b.mat, nl);
  mat_mulvec(o->pos, b.mat, pl); v3add(o->pos, o->pos, b.pos);
  return 1;
}
// minimiser interval of dist^2(p0 + t*dir, box) over t in [0,1]; exact piecewise-quadratic scan
DEV void seg_box_min(const real* p0, const real* dir, const real* size, real* tlo, real* thi) {
  real bp[8]; int nb = 0;
  bp[nb++] = 0; bp[nb++] = 1;
  for (int i = 0; i < 3; i++)
    if (fabs(dir[i]) > (real)1e-9)
      for (int s = -1; s <= 1; s += 2) { real t = (s * size[i] - p0[i]) / dir[i]; if (t > 0 && t < 1) bp[nb++] = t; }
  for (int i = 1; i < nb; i++) { real x = bp[i]; int j = i - 1; while (j >= 0 && bp[j] > x) { bp[j + 1] = bp[j]; j--; } bp[j + 1] = x; }
  real best = (real)1e30, blo = 0, bhi = 0;
  for (int k = 0; k + 1 < nb; k++) {
    real t0 = bp[k], t1 = bp[k + 1], tm = (real)0.5 * (t0 + t1);
    real A = 0, B = 0, C = 0;
    for (int i = 0; i < 3; i++) {
      real x = p0[i] + tm * dir[i];
      if (x > size[i] || x < -size[i]) { real aa = p0[i] + (x > size[i] ? -size[i] : size[i]), bb = dir[i]; A += bb * bb; B += 2 * aa * bb; C += aa * aa; }
    }
    real tl, th, f;
    real f0 = A * t0 * t0 + B * t0 + C, f1 = A * t1 * t1 + B * t1 + C;
    real ts = A > 0 ? fmin(fmax(-B / (2 * A), t0), t1) : t0;
    f = A * ts * ts + B * ts + C;
    const real rel = sizeof(real) == 4 ? (real)1e-5 : (real)1e-9;
    if (fmax(f0, f1) - f <= rel * f + (real)1e-18) { tl = t0; th = t1; f = fmin(fmin(f0, f1), f); }   // flat: segment parallel to the face
    else { tl = th = ts; }
    real tol = rel * best + (real)1e-18;
    if (f < best - tol) { best = f; blo = tl; bhi = th; }
    else if (fabs(f - best) <= tol && tl <= bhi + (real)1e-6) { if (th > bhi) bhi = th; }
  }
  *tlo = blo; *thi = bhi;
}
DEV int capsule_box(const DShape& a, const DShape& b, real margin, RawCon* o) {
  real ax[3], t[3], p0[3], p1[3], l0[3], l1[3], dir[3];
  mat_col(ax, a.mat, 2);
  v3addscl(p0, a.pos, ax, -a.size[1]); v3addscl(p1, a.pos, ax, a.size[1]);
  v3sub(t, p0, b.pos); mat_tmulvec(l0, b.mat, t);
  v3sub(t, p1, b.pos); mat_tmulvec(l1, b.mat, t);
  v3sub(dir, l1, l0);
  real tlo, thi;
  seg_box_min(l0, dir, b.size, &tlo, &thi);
  int cnt = 0; real c[3];
  v3addscl(c, p0, ax, 2 * a.size[1] * tlo); cnt += sphere_box_raw(c, a.size[0], b, margin, o + cnt);
  if (thi - tlo > (real)1e-5) { v3addscl(c, p0, ax, 2 * a.size[1] * thi); cnt += sphere_box_raw(c, a.size[0], b, margin, o + cnt); }
  return cnt;
}

// ------------------------------------------------------------------ box-box: separating axes + face clipping
DEV int clip_poly(real* poly, int n, int axis, real lim, real sgn) {
  real out[32]; int m = 0;
  for (int i = 0; i < n; i++) {
    const real* p = poly + 2 * i; const real* q = poly + 2 * ((i + 1) % n);
    real dp = sgn * p[axis] - lim, dq = sgn * q[axis] - lim;
    if (dp <= 0) { out[2 * m] = p[0]; out[2 * m + 1] = p[1]; m++; }
    if ((dp < 0 && dq > 0) || (dp > 0 && dq < 0)) {
      real s = dp / (dp - dq);
      out[2 * m] = p[0] + s * (q[0] - p[0]); out[2 * m + 1] = p[1] + s * (q[1] - p[1]); m++;
    }
    if (m >= 15) break;
  }
  for (int i = 0; i < 2 * m; i++) poly[i] = out[i];
  return m;
}
__device__ __noinline__ int box_box(const DShape& a, const DShape& b, real margin, RawCon* o) {
  const real* R1 = a.mat; const real* R2 = b.mat;
  real p[3], pp[3]; v3sub(p, b.pos, a.pos); mat_tmulvec(pp, R1, p);
  real R[9], Q[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    real c1[3], c2[3]; mat_col(c1, R1, i); mat_col(c2, R2, j);
    R[3 * i + j] = v3dot(c1, c2); Q[3 * i + j] = fabs(R[3 * i + j]);
  }
  const real* A = a.size; const real* B = b.size;
  real s = (real)-1e30; int code = 0; bool invert = false, haveC = false; real normalC[3] = {0, 0, 0};
  for (int i = 0; i < 3; i++) {
    real e = fabs(pp[i]) - (A[i] + B[0] * Q[3 * i] + B[1] * Q[3 * i + 1] + B[2] * Q[3 * i + 2]);
    if (e > margin) return 0;
    if (e > s) { s = e; code = i + 1; invert = pp[i] < 0; haveC = false; }
  }
  for (int j = 0; j < 3; j++) {
    real c2[3]; mat_col(c2, R2, j);
    real e1 = v3dot(c2, p);
    real e = fabs(e1) - (A[0] * Q[j] + A[1] * Q[3 + j] + A[2] * Q[6 + j] + B[j]);
    if (e > margin) return 0;
    if (e > s) { s = e; code = j + 4; invert = e1 < 0; haveC = false; }
  }
  const real fudge = (real)1.05;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    real n[3] = {0, 0, 0};
    n[i1] = -R[3 * i2 + j]; n[i2] = R[3 * i1 + j];
    real l = sqrt(n[i1] * n[i1] + n[i2] * n[i2]);
    if (l < (real)1e-5) continue;
    real e1 = pp[i2] * R[3 * i1 + j] - pp[i1] * R[3 * i2 + j];
    real e = fabs(e1) - (A[i1] * Q[3 * i2 + j] + A[i2] * Q[3 * i1 + j] + B[j1] * Q[3 * i + j2] + B[j2] * Q[3 * i + j1]);
    e /= l;
    if (e > margin) return 0;
    if ((e < 0 ? e * fudge : e) > s) {
      s = e; code = 7 + 3 * i + j; invert = e1 < 0; haveC = true;
      normalC[0] = n[0] / l; normalC[1] = n[1] / l; normalC[2] = n[2] / l;
    }
  }
  if (!code) return 0;
  real normal[3];
  if (haveC) mat_mulvec(normal, R1, normalC);
  else if (code <= 3) mat_col(normal, R1, code - 1);
  else mat_col(normal, R2, code - 4);
  if (invert) v3scl(normal, normal, -1);
  real depth = -s;
  if (code > 6) {
    real pa[3], pb[3]; v3copy(pa, a.pos); v3copy(pb, b.pos);
    for (int j = 0; j < 3; j++) {
      real c1[3], c2[3]; mat_col(c1, R1, j); mat_col(c2, R2, j);
      real sg = v3dot(normal, c1) > 0 ? (real)1 : (real)-1; v3addscl(pa, pa, c1, sg * A[j]);
      sg = v3dot(normal, c2) > 0 ? (real)-1 : (real)1; v3addscl(pb, pb, c2, sg * B[j]);
    }
    int ia = (code - 7) / 3, ib = (code - 7) % 3;
    real ua[3], ub[3]; mat_col(ua, R1, ia); mat_col(ub, R2, ib);
    real d[3]; v3sub(d, pb, pa);
    real uaub = v3dot(ua, ub), q1 = v3dot(ua, d), q2 = -v3dot(ub, d), dd = 1 - uaub * uaub;
    real alpha = 0, beta = 0;
    if (dd > (real)1e-4) { alpha = (q1 + uaub * q2) / dd; beta = (uaub * q1 + q2) / dd; }
    v3addscl(pa, pa, ua, alpha); v3addscl(pb, pb, ub, beta);
    o->dist = -depth; v3copy(o->normal, normal);
    for (int k = 0; k < 3; k++) o->pos[k] = (real)0.5 * (pa[k] + pb[k]);
    return 1;
  }
  const real *Ra, *Rb, *pa, *pb, *Sa, *Sb; real nrm[3];
  if (code <= 3) { Ra = R1; Rb = R2; pa = a.pos; pb = b.pos; Sa = A; Sb = B; v3copy(nrm, normal); }
  else { Ra = R2; Rb = R1; pa = b.pos; pb = a.pos; Sa = B; Sb = A; v3scl(nrm, normal, -1); }
  real nr[3], anr[3]; mat_tmulvec(nr, Rb, nrm);
  for (int k = 0; k < 3; k++) anr[k] = fabs(nr[k]);
  int lanr, a1, a2;
  if (anr[1] > anr[0]) { if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; } }
  else { if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; } }
  real center[3], col[3]; mat_col(col, Rb, lanr);
  real sg = nr[lanr] < 0 ? (real)1 : (real)-1;
  for (int k = 0; k < 3; k++) center[k] = pb[k] - pa[k] + sg * Sb[lanr] * col[k];
  int codeN = (code <= 3 ? code - 1 : code - 4), code1, code2;
  if (codeN == 0) { code1 = 1; code2 = 2; } else if (codeN == 1) { code1 = 0; code2 = 2; } else { code1 = 0; code2 = 1; }
  real r1[3], r2[3], i1v[3], i2v[3];
  mat_col(r1, Ra, code1); mat_col(r2, Ra, code2); mat_col(i1v, Rb, a1); mat_col(i2v, Rb, a2);
  real c1 = v3dot(center, r1), c2 = v3dot(center, r2);
  real m11 = v3dot(r1, i1v), m12 = v3dot(r1, i2v), m21 = v3dot(r2, i1v), m22 = v3dot(r2, i2v);
  real k1 = m11 * Sb[a1], k2 = m21 * Sb[a1], k3 = m12 * Sb[a2], k4 = m22 * Sb[a2];
  real quad[32];
  quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4; quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
  quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4; quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
  int n = 4;
  n = clip_poly(quad, n, 0, Sa[code1], 1); if (n) n = clip_poly(quad, n, 0, Sa[code1], -1);
  if (n) n = clip_poly(quad, n, 1, Sa[code2], 1); if (n) n = clip_poly(quad, n, 1, Sa[code2], -1);
  if (n < 1) return 0;
  real det1 = 1 / (m11 * m22 - m12 * m21);
  m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
  int cnt = 0;
  for (int j = 0; j < n && cnt < 8; j++) {
    real kk1 = m22 * (quad[2 * j] - c1) - m12 * (quad[2 * j + 1] - c2);
    real kk2 = -m21 * (quad[2 * j] - c1) + m11 * (quad[2 * j + 1] - c2);
    real pt[3];
    for (int k = 0; k < 3; k++) pt[k] = center[k] + kk1 * i1v[k] + kk2 * i2v[k];
    real dep = Sa[codeN] - v3dot(nrm, pt);
    if (dep < -margin) continue;
    real cp[3];
    for (int k = 0; k < 3; k++) cp[k] = pt[k] + pa[k] + (real)0.5 * dep * nrm[k];
    bool dup = false;
    for (int q = 0; q < cnt; q++) { real dx[3]; v3sub(dx, o[q].pos, cp); if (v3dot(dx, dx) < (real)1e-12) dup = true; }
    if (dup) continue;
    o[cnt].dist = -dep; v3copy(o[cnt].normal, normal); v3copy(o[cnt].pos, cp);
    cnt++;
  }
  return cnt;
}

// one analytic pair on the calling lane; returns #contacts (<= 8)
DEV int narrow_analytic(const DShape& a, const DShape& b, real margin, RawCon* o) {
  int t1 = a.type, t2 = b.type;
  if (t1 == G_PLANE) {
    real n[3]; mat_col(n, a.mat, 2);
    if (t2 == G_SPHERE) return plane_sphere_raw(n, a.pos, b.pos, b.size[0], margin, o);
    if (t2 == G_CAPSULE) return plane_capsule(a, b, margin, o);
    if (t2 == G_CYLINDER) return plane_cylinder(a, b, margin, o);
    if (t2 == G_BOX) return plane_box(a, b, margin, o);
    return 0;
  }
  if (t1 == G_SPHERE && t2 == G_SPHERE) return sphere_sphere_raw(a.pos, a.size[0], b.pos, b.size[0], margin, o);
  if (t1 == G_SPHERE && t2 == G_CAPSULE) return sphere_capsule(a, b, margin, o);
  if (t1 == G_SPHERE && t2 == G_BOX) return sphere_box_raw(a.pos, a.size[0], b, margin, o);
  if (t1 == G_CAPSULE && t2 == G_CAPSULE) return capsule_capsule(a, b, margin, o);
  if (t1 == G_CAPSULE && t2 == G_BOX) return capsule_box(a, b, margin, o);
  if (t1 == G_BOX && t2 == G_BOX) return box_box(a, b, margin, o);
  return 0;
}
DEV bool pair_is_analytic(int t1, int t2) {
  if (t1 == G_PLANE) return t2 != G_MESH;
  if (t1 == G_SPHERE) return t2 == G_SPHERE || t2 == G_CAPSULE || t2 == G_BOX;
  if (t1 == G_CAPSULE) return t2 == G_CAPSULE || t2 == G_BOX;
  return t1 == G_BOX && t2 == G_BOX;
}

// ------------------------------------------------------------------ general convex: GJK + EPA, warp-cooperative
struct SV { real v[3], a[3], b[3]; };

// support point of one shape for a world direction; evaluated by the whole warp (uniform result)
DEV void support_shape(const DShape& s, const real* dir, real* out, int lane) {
  real dl[3], l[3];
  mat_tmulvec(dl, s.mat, dir);
  if (s.type == G_MESH) {
    real bv = (real)-1e30; int bi = 0;
    for (int i = lane; i < s.nvert; i += 32) {
      real x = dl[0] * __ldg(s.vert + 3 * i) + dl[1] * __ldg(s.vert + 3 * i + 1) + dl[2] * __ldg(s.vert + 3 * i + 2);
      if (x > bv) { bv = x; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      real ov = __shfl_xor_sync(FULLMASK, bv, o); int oi = __shfl_xor_sync(FULLMASK, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    l[0] = __ldg(s.vert + 3 * bi); l[1] = __ldg(s.vert + 3 * bi + 1); l[2] = __ldg(s.vert + 3 * bi + 2);
  } else if (s.type == G_BOX) {
    for (int i = 0; i < 3; i++) l[i] = dl[i] >= 0 ? s.size[i] : -s.size[i];
  } else if (s.type == G_CYLINDER) {
    real n = sqrt(dl[0] * dl[0] + dl[1] * dl[1]);
    if (n > MW_EPS) { l[0] = dl[0] * s.size[0] / n; l[1] = dl[1] * s.size[0] / n; } else l[0] = l[1] = 0;
    l[2] = dl[2] >= 0 ? s.size[1] : -s.size[1];
  } else if (s.type == G_CAPSULE) {
    l[0] = l[1] = 0; l[2] = dl[2] >= 0 ? s.size[1] : -s.size[1];
  } else {
    l[0] = l[1] = l[2] = 0;
  }
  mat_mulvec(out, s.mat, l); v3add(out, out, s.pos);
}
DEV real core_radius(const DShape& s) { return (s.type == G_SPHERE || s.type == G_CAPSULE) ? s.size[0] : (real)0; }
DEV void support_pair(const DShape& A, const DShape& B, const real* dir, SV* o, int lane) {
  real nd[3] = {-dir[0], -dir[1], -dir[2]};
  support_shape(A, dir, o->a, lane);
  support_shape(B, nd, o->b, lane);
  v3sub(o->v, o->a, o->b);
}
DEV void closest_tri(const real* a, const real* b, const real* c, real* w) {
  real ab[3], ac[3], ap[3]; v3sub(ab, b, a); v3sub(ac, c, a); v3scl(ap, a, -1);
  real d1 = v3dot(ab, ap), d2 = v3dot(ac, ap);
  if (d1 <= 0 && d2 <= 0) { w[0] = 1; w[1] = w[2] = 0; return; }
  real bp[3]; v3scl(bp, b, -1);
  real d3 = v3dot(ab, bp), d4 = v3dot(ac, bp);
  if (d3 >= 0 && d4 <= d3) { w[1] = 1; w[0] = w[2] = 0; return; }
  real vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { real v = d1 / (d1 - d3); w[0] = 1 - v; w[1] = v; w[2] = 0; return; }
  real cp[3]; v3scl(cp, c, -1);
  real d5 = v3dot(ab, cp), d6 = v3dot(ac, cp);
  if (d6 >= 0 && d5 <= d6) { w[2] = 1; w[0] = w[1] = 0; return; }
  real vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { real x = d2 / (d2 - d6); w[0] = 1 - x; w[1] = 0; w[2] = x; return; }
  real va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) { real x = (d4 - d3) / ((d4 - d3) + (d5 - d6)); w[0] = 0; w[1] = 1 - x; w[2] = x; return; }
  real den = 1 / (va + vb + vc);
  w[1] = vb * den; w[2] = vc * den; w[0] = 1 - w[1] - w[2];
}
// simplex reduction (leader lane only). returns true if the origin is enclosed by a tetrahedron
__device__ __noinline__ bool closest_simplex(SV* s, int* n, real* v) {
  real w[4] = {0, 0, 0, 0};
  if (*n == 1) w[0] = 1;
  else if (*n == 2) {
    real ab[3]; v3sub(ab, s[1].v, s[0].v);
    real t = -v3dot(s[0].v, ab) / fmax(v3dot(ab, ab), (real)1e-30);
    if (t <= 0) w[0] = 1; else if (t >= 1) w[1] = 1; else { w[0] = 1 - t; w[1] = t; }
  } else if (*n == 3) closest_tri(s[0].v, s[1].v, s[2].v, w);
  else {
    const int F[4][4] = {{0, 1, 2, 3}, {0, 2, 3, 1}, {0, 3, 1, 2}, {1, 3, 2, 0}};
    real best = (real)1e30; bool found = false; real bw[4] = {0, 0, 0, 0};
    for (int f = 0; f < 4; f++) {
      const real *a = s[F[f][0]].v, *b = s[F[f][1]].v, *c = s[F[f][2]].v, *dv = s[F[f][3]].v;
      real ab[3], ac[3], nrm[3], ad[3]; v3sub(ab, b, a); v3sub(ac, c, a); v3cross(nrm, ab, ac); v3sub(ad, dv, a);
      real sd = v3dot(nrm, ad), so = -v3dot(nrm, a);
      if (fabs(sd) < (real)1e-30) { so = 1; sd = -1; }
      if ((sd > 0 && so > 0) || (sd < 0 && so < 0)) continue;
      real tw[3]; closest_tri(a, b, c, tw);
      real q[3]; for (int k = 0; k < 3; k++) q[k] = tw[0] * a[k] + tw[1] * b[k] + tw[2] * c[k];
      real dd = v3dot(q, q);
      if (dd < best) { best = dd; found = true; bw[0] = bw[1] = bw[2] = bw[3] = 0; bw[F[f][0]] = tw[0]; bw[F[f][1]] = tw[1]; bw[F[f][2]] = tw[2]; }
    }
    if (!found) { v3zero(v); return true; }
    for (int i = 0; i < 4; i++) w[i] = bw[i];
  }
  int m = 0; v3zero(v);
  for (int i = 0; i < *n; i++) if (w[i] > 0) { for (int k = 0; k < 3; k++) v[k] += w[i] * s[i].v[k]; if (m != i) s[m] = s[i]; m++; }
  *n = m;
  return false;
}
DEV void simplex_weights(const SV* s, int n, real* w) {
  w[0] = 1; w[1] = w[2] = 0;
  if (n == 2) {
    real ab[3]; v3sub(ab, s[1].v, s[0].v);
    real t = fmin(fmax(-v3dot(s[0].v, ab) / fmax(v3dot(ab, ab), (real)1e-30), (real)0), (real)1);
    w[0] = 1 - t; w[1] = t;
  } else if (n == 3) closest_tri(s[0].v, s[1].v, s[2].v, w);
}

#define EPA_MAXV 80
#define EPA_MAXF 160
#define EPA_MAXE 96
struct EFace { int v[3]; real n[3], d; int alive; };
struct EpaWs { SV V[EPA_MAXV]; EFace F[EPA_MAXF]; int edge[EPA_MAXE][2]; };
#define EPA_WS_WORDS ((int)(sizeof(EpaWs) / 4))

DEV bool epa_add_face(EpaWs* W, int* nf, int a, int b, int c) {
  if (*nf >= EPA_MAXF) return false;
  EFace* f = W->F + (*nf);
  f->v[0] = a; f->v[1] = b; f->v[2] = c; f->alive = 1;
  real ab[3], ac[3], nn[3]; v3sub(ab, W->V[b].v, W->V[a].v); v3sub(ac, W->V[c].v, W->V[a].v); v3cross(nn, ab, ac);
  real l = v3norm(nn);
  if (l < (real)1e-30) { f->d = 0; nn[0] = 1; nn[1] = nn[2] = 0; }
  else { v3scl(nn, nn, 1 / l); f->d = v3dot(nn, W->V[a].v); }
  if (f->d < 0) { int t = f->v[1]; f->v[1] = f->v[2]; f->v[2] = t; v3scl(nn, nn, -1); f->d = -f->d; }
  v3copy(f->n, nn);
  (*nf)++;
  return true;
}


// Face-on contact between a cylinder and a box face (cap-on-face or generator-line-on-face): every point of the
// contact patch is a valid EPA witness, so the contact *point* is implementation-defined in the reference physics.
// We take the pressure centroid of the patch under a linear penetration profile, clipped to the overlap with the
// box face (continuous in the pose; removes rim-to-rim chatter of resting cylinders).  Other configurations are
// left as EPA/GJK produced them.
__device__ __noinline__ void refine_cyl_box(const DShape& A, const DShape& B, RawCon* o) {
  if (A.type != G_CYLINDER || B.type != G_BOX) return;
  real nb[3]; mat_tmulvec(nb, B.mat, o->normal);
  int k = 0; for (int q = 1; q < 3; q++) if (fabs(nb[q]) > fabs(nb[k])) k = q;
  if (fabs(nb[k]) < (real)0.99999) return;                       /* the contact normal must be a face normal of the box */
  const real sgn = nb[k] > 0 ? (real)1 : (real)-1;
  const int i = (k + 1) % 3, j = (k + 2) % 3;
  const real r = A.size[0], h = A.size[1];
  real ax[3], a[3], t[3], cB[3];
  mat_col(ax, A.mat, 2); mat_tmulvec(a, B.mat, ax);
  v3sub(t, A.pos, B.pos); mat_tmulvec(cB, B.mat, t);
  const real ak = a[k] * sgn;
  real P[3], pen;
  if (fabs(ak) > (real)0.9) {
    /* ---- cap on face */
    const real cs = ak > 0 ? (real)1 : (real)-1;
    real C[3], acap[3], g[3];
    for (int q = 0; q < 3; q++) { acap[q] = cs * a[q]; C[q] = cB[q] + cs * h * a[q]; }
    const real pen0 = sgn * C[k] + B.size[k];
    for (int q = 0; q < 3; q++) g[q] = -(sgn * acap[k]) * acap[q];
    g[k] += sgn;
    const real st = v3norm(g);                                   /* sin(tilt) */
    real sbar = 0;
    if (st > (real)1e-9) sbar = pen0 > 0 ? fmin(r, r * r * st / (4 * pen0)) : r;
    for (int q = 0; q < 3; q++) P[q] = C[q] + (st > (real)1e-9 ? g[q] / st * sbar : (real)0);
    const int axs[2] = {i, j};
    for (int q = 0; q < 2; q++) {
      const int u = axs[q];
      const real lo = fmax(C[u] - r, -B.size[u]), hi = fmin(C[u] + r, B.size[u]);
      if (lo > hi) return;
      P[u] = fmin(fmax(P[u], lo), hi);
    }
    P[k] = C[k] - (acap[i] * (P[i] - C[i]) + acap[j] * (P[j] - C[j])) / acap[k];
    pen = sgn * P[k] + B.size[k];
  } else if (fabs(ak) < (real)0.1) {
    /* ---- generator line on face */
    real d[3], L0[3];
    for (int q = 0; q < 3; q++) d[q] = -(sgn * a[k]) * a[q];
    d[k] += sgn;
    v3normalize(d);
    for (int q = 0; q < 3; q++) L0[q] = cB[q] + r * d[q];
    real t0 = -h, t1 = h;
    const int axs[2] = {i, j};
    for (int q = 0; q < 2; q++) {
      const int u = axs[q];
      if (fabs(a[u]) > (real)1e-9) {
        const real ta = (-B.size[u] - L0[u]) / a[u], tb = (B.size[u] - L0[u]) / a[u];
        t0 = fmax(t0, fmin(ta, tb)); t1 = fmin(t1, fmax(ta, tb));
      } else if (fabs(L0[u]) > B.size[u]) return;
    }
    if (t0 > t1) return;
    const real tm = (real)0.5 * (t0 + t1), Lh = (real)0.5 * (t1 - t0), sl = sgn * a[k];
    const real penm = sgn * (L0[k] + tm * a[k]) + B.size[k];
    real off;
    if (penm > 0) off = fmin(fmax(sl * Lh * Lh / (3 * penm), -Lh), Lh);
    else off = sl > 0 ? Lh : (sl < 0 ? -Lh : (real)0);
    const real ts = tm + off;
    for (int q = 0; q < 3; q++) P[q] = L0[q] + ts * a[q];
    pen = sgn * P[k] + B.size[k];
  } else return;
  o->dist = -pen;
  P[k] -= sgn * (real)0.5 * pen;
  mat_mulvec(o->pos, B.mat, P); v3add(o->pos, o->pos, B.pos);
}

// Whole warp calls this with identical A, B.  The result (count 0/1, contact in *o) is valid on every lane.
__device__ __noinline__ int convex_pair(const DShape& A, const DShape& B, real margin, RawCon* o, EpaWs* W, int lane) {
  const real dirs[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  real ra = core_radius(A), rb = core_radius(B);
  // ---- leader state
  SV s[4]; int n = 0; real v[3]; v3sub(v, A.pos, B.pos);
  if (v3dot(v, v) < (real)1e-20) { v[0] = 1; v[1] = v[2] = 0; }
  int status = 0;   // 0 running, 1 closest point found (separated), 2 enclosed -> EPA, 3 no contact
  const real reltol = sizeof(real) == 4 ? (real)1e-6 : (real)1e-12;
  for (int it = 0; it < 48 && status == 0; it++) {
    real dir[3] = {bcast(-v[0], 0), bcast(-v[1], 0), bcast(-v[2], 0)};
    SV w; support_pair(A, B, dir, &w, lane);
    if (lane == 0) {
      if (n == 0) { s[0] = w; n = 1; v3copy(v, w.v); if (v3dot(v, v) < (real)1e-20) status = 2; }
      else {
        real vv = v3dot(v, v), vw = v3dot(v, w.v);
        if (vv - vw <= reltol * vv) status = 1;
        else if (vw > 0 && vw / sqrt(vv) - ra - rb > margin + (real)1e-4) status = 3;
        else {
          bool dup = false;
          for (int i = 0; i < n; i++) { real t[3]; v3sub(t, s[i].v, w.v); if (v3dot(t, t) < (real)1e-20) dup = true; }
          if (dup) status = 1;
          else {
            s[n++] = w;
            if (closest_simplex(s, &n, v)) status = 2;
            else if (v3dot(v, v) < (real)1e-20) status = 2;
          }
        }
      }
    }
    status = __shfl_sync(FULLMASK, status, 0);
  }
  if (status == 0) status = 1;
  int result = 0;
  RawCon rc; rc.dist = 0; v3zero(rc.pos); v3zero(rc.normal);
  if (status == 1 && lane == 0) {
    real w[3]; simplex_weights(s, n, w);
    real wa[3] = {0, 0, 0}, wb[3] = {0, 0, 0};
    for (int i = 0; i < n; i++) { v3addscl(wa, wa, s[i].a, w[i]); v3addscl(wb, wb, s[i].b, w[i]); }
    real dvec[3]; v3sub(dvec, wb, wa);
    real dcore = v3norm(dvec);
    if (dcore > (real)1e-7) {
      real dist = dcore - ra - rb;
      if (dist <= margin + (real)1e-4) {      // slack: the analytic refinement below makes the final call
        v3scl(rc.normal, dvec, 1 / dcore); rc.dist = dist;
        for (int k = 0; k < 3; k++) rc.pos[k] = (real)0.5 * (wa[k] + rc.normal[k] * ra + wb[k] - rc.normal[k] * rb);
        refine_cyl_box(A, B, &rc);
        result = rc.dist <= margin ? 1 : -1;
      } else result = -1;
    } else status = 2;   // touching: let EPA resolve the direction
    if (result == -1) { result = 0; status = 3; }
  }
  status = __shfl_sync(FULLMASK, status, 0);
  if (status == 2) {
    // ---- EPA.  phase 0: grow the simplex to a tetrahedron; phase 1: expand the polytope
    int nv = 0, nf = 0, k = 0, phase = 0, done = 0, bestf = -1;   // leader state
    real dirl[3] = {1, 0, 0};
    for (int it = 0; it < 176; it++) {
      if (lane == 0) {
        // choose the next query direction
        if (phase == 0) {
          if (n == 1) { if (k >= 6) done = 2; else v3copy(dirl, dirs[k]); }
          else if (n == 2) {
            real ab[3]; v3sub(ab, s[1].v, s[0].v);
            bool ok = false;
            while (k < 12 && !ok) {
              v3cross(dirl, ab, dirs[k >> 1]);
              if (v3dot(dirl, dirl) >= (real)1e-8 * v3dot(ab, ab)) ok = true; else k++;
            }
            if (!ok) done = 2; else if (k & 1) v3scl(dirl, dirl, -1);
          } else if (n == 3) {
            if (k >= 2) done = 2;
            else { real ab[3], ac[3]; v3sub(ab, s[1].v, s[0].v); v3sub(ac, s[2].v, s[0].v); v3cross(dirl, ab, ac); if (k) v3scl(dirl, dirl, -1); }
          }
          if (n == 4) {
            for (int i = 0; i < 4; i++) W->V[i] = s[i];
            nv = 4; nf = 0;
            epa_add_face(W, &nf, 0, 1, 2); epa_add_face(W, &nf, 0, 2, 3); epa_add_face(W, &nf, 0, 3, 1); epa_add_face(W, &nf, 1, 3, 2);
            phase = 1;
          }
        }
        if (phase == 1 && !done) {
          bestf = -1; real bd = (real)1e30;
          for (int f = 0; f < nf; f++) if (W->F[f].alive && W->F[f].d < bd) { bd = W->F[f].d; bestf = f; }
          if (bestf < 0) done = 2; else v3copy(dirl, W->F[bestf].n);
        }
      }
      done = __shfl_sync(FULLMASK, done, 0);
      if (done) break;
      real dir[3] = {bcast(dirl[0], 0), bcast(dirl[1], 0), bcast(dirl[2], 0)};
      SV w; support_pair(A, B, dir, &w, lane);
      if (lane == 0) {
        if (phase == 0) {
          if (n == 1) { real dd[3]; v3sub(dd, w.v, s[0].v); if (v3dot(dd, dd) > (real)1e-14) { s[n++] = w; k = 0; } else k++; }
          else if (n == 2) {
            real ab[3], aw[3], cr[3]; v3sub(ab, s[1].v, s[0].v); v3sub(aw, w.v, s[0].v); v3cross(cr, ab, aw);
            if (v3dot(cr, cr) > (real)1e-14 * v3dot(ab, ab)) { s[n++] = w; k = 0; } else k++;
          } else if (n == 3) {
            real ab[3], ac[3], nn[3], aw[3]; v3sub(ab, s[1].v, s[0].v); v3sub(ac, s[2].v, s[0].v); v3cross(nn, ab, ac); v3sub(aw, w.v, s[0].v);
            if (fabs(v3dot(aw, nn)) > (real)1e-7 * sqrt(v3dot(nn, nn))) { s[n++] = w; k = 0; } else k++;
          }
        } else {
          real bd = W->F[bestf].d;
          real dw = v3dot(w.v, W->F[bestf].n);
          real tol = sizeof(real) == 4 ? (real)1e-6 : (real)1e-10;
          if (dw - bd < tol || nv >= EPA_MAXV) done = 1;
          else {
            int ne = 0; bool ovf = false;
            for (int f = 0; f < nf; f++) {
              if (!W->F[f].alive) continue;
              real t[3]; v3sub(t, w.v, W->V[W->F[f].v[0]].v);
              if (v3dot(W->F[f].n, t) > (real)1e-9) {
                W->F[f].alive = 0;
                for (int e = 0; e < 3; e++) {
                  int ea = W->F[f].v[e], eb = W->F[f].v[(e + 1) % 3], found = -1;
                  for (int q = 0; q < ne; q++) if (W->edge[q][0] == eb && W->edge[q][1] == ea) { found = q; break; }
                  if (found >= 0) { W->edge[found][0] = W->edge[ne - 1][0]; W->edge[found][1] = W->edge[ne - 1][1]; ne--; }
                  else if (ne < EPA_MAXE) { W->edge[ne][0] = ea; W->edge[ne][1] = eb; ne++; }
                  else ovf = true;
                }
              }
            }
            if (ne == 0 || ovf) done = 1;
            else {
              int wi = nv; W->V[nv++] = w;
              for (int q = 0; q < ne; q++) if (!epa_add_face(W, &nf, W->edge[q][0], W->edge[q][1], wi)) { done = 1; break; }
            }
          }
        }
      }
    }
    if (lane == 0 && done != 2 && phase == 1 && bestf >= 0) {
      const EFace* f = W->F + bestf;
      real w3[3]; closest_tri(W->V[f->v[0]].v, W->V[f->v[1]].v, W->V[f->v[2]].v, w3);
      real wa[3] = {0, 0, 0}, wb[3] = {0, 0, 0};
      for (int q = 0; q < 3; q++) { v3addscl(wa, wa, W->V[f->v[q]].a, w3[q]); v3addscl(wb, wb, W->V[f->v[q]].b, w3[q]); }
      real dist = -f->d - ra - rb;
      if (dist <= margin) {
        rc.dist = dist; v3copy(rc.normal, f->n);
        for (int q = 0; q < 3; q++) rc.pos[q] = (real)0.5 * (wa[q] + rc.normal[q] * ra + wb[q] - rc.normal[q] * rb);
        refine_cyl_box(A, B, &rc);
        result = rc.dist <= margin ? 1 : 0;
      }
    }
  }
  result = __shfl_sync(FULLMASK, result, 0);
  o->dist = bcast(rc.dist, 0);
  for (int q = 0; q < 3; q++) { o->pos[q] = bcast(rc.pos[q], 0); o->normal[q] = bcast(rc.normal[q], 0); }
  return result;
}
