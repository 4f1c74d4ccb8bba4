/* oracle/mjcollide.c -- TEST INFRASTRUCTURE (see mjphys.h: PARITY UNPINNED).
 *
 * Collision detection for the CPU oracle: restates what MuJoCo's mj_collision
 * [3P] does for the geom types Meta-World uses.  Static pair filtering is done
 * once in om_model_finalize; here: bounding-sphere cull, narrowphase, contact
 * parameter mixing.  Convention (MuJoCo's): geom1 has the lower type id, the
 * normal points from geom1 to geom2, `dist` < 0 is penetration, `pos` is the
 * midpoint between the two surfaces.
 *
 * Analytic pairs: plane-{sphere,capsule,cylinder,box,mesh}, sphere-sphere,
 * sphere-capsule, capsule-capsule, sphere-box, capsule-box, box-box.
 * Everything else (cylinder and mesh pairs) goes through one general convex
 * routine (GJK distance + EPA penetration, one contact per pair = MuJoCo with
 * multiccd off).
 */
#include <stdlib.h>
#include "mjinternal.h"

typedef struct { double dist, pos[3], normal[3]; } RawCon;

typedef struct {
  int type; const double* pos; const double* mat; const double* size; const double* vert; int nvert;
} Shape;

static void get_shape(const OModel* m, const OData* d, int g, Shape* s) {
  s->type = m->geom_type[g]; s->pos = d->geom_xpos + 3 * g; s->mat = d->geom_xmat + 9 * g; s->size = m->geom_size + 3 * g;
  s->vert = NULL; s->nvert = 0;
  if (s->type == G_MESH) { int id = m->geom_dataid[g]; s->vert = m->mesh_vert[id]; s->nvert = m->mesh_nvert[id]; }
}
static void mat_col(double* r, const double* M, int c) { r[0] = M[c]; r[1] = M[3 + c]; r[2] = M[6 + c]; }

/* ------------------------------------------------------------------ plane pairs */
static int plane_sphere_raw(const double* pn, const double* pp, const double* c, double r, double margin, RawCon* o) {
  double t[3]; v3sub(t, c, pp);
  double dist = v3dot(t, pn) - r;
  if (dist > margin) return 0;
  o->dist = dist; v3copy(o->normal, pn);
  v3addscl(o->pos, c, pn, -(r + 0.5 * dist));
  return 1;
}
static int plane_sphere(const Shape* a, const Shape* b, double margin, RawCon* o) {
  double n[3]; mat_col(n, a->mat, 2);
  return plane_sphere_raw(n, a->pos, b->pos, b->size[0], margin, o);
}
static int plane_capsule(const Shape* a, const Shape* b, double margin, RawCon* o) {
  double n[3], ax[3], c[3]; mat_col(n, a->mat, 2); mat_col(ax, b->mat, 2);
  int cnt = 0;
  v3addscl(c, b->pos, ax, b->size[1]); cnt += plane_sphere_raw(n, a->pos, c, b->size[0], margin, o + cnt);
  v3addscl(c, b->pos, ax, -b->size[1]); cnt += plane_sphere_raw(n, a->pos, c, b->size[0], margin, o + cnt);
  return cnt;
}
static int plane_cylinder(const Shape* a, const Shape* b, double margin, RawCon* o) {
  double n[3], axis[3], vec[3], t[3];
  mat_col(n, a->mat, 2); mat_col(axis, b->mat, 2);
  v3sub(t, b->pos, a->pos);
  double dist0 = v3dot(t, n), prjaxis = v3dot(n, axis);
  if (prjaxis > 0) { v3scl(axis, axis, -1); prjaxis = -prjaxis; }
  for (int i = 0; i < 3; i++) vec[i] = prjaxis * axis[i] - n[i];
  double len = v3norm(vec);
  if (len >= 1e-12) v3scl(vec, vec, b->size[0] / len);
  else { mat_col(vec, b->mat, 0); v3scl(vec, vec, b->size[0]); }
  double prjvec = v3dot(vec, n);
  v3scl(axis, axis, b->size[1]); prjaxis *= b->size[1];
  int cnt = 0;
  if (dist0 + prjaxis + prjvec > margin) return 0;
  double dd = dist0 + prjaxis + prjvec;
  o[cnt].dist = dd; v3copy(o[cnt].normal, n);
  for (int i = 0; i < 3; i++) o[cnt].pos[i] = b->pos[i] + vec[i] + axis[i] - n[i] * dd * 0.5;
  cnt++;
  if (dist0 - prjaxis + prjvec <= margin) {
    dd = dist0 - prjaxis + prjvec;
    o[cnt].dist = dd; v3copy(o[cnt].normal, n);
    for (int i = 0; i < 3; i++) o[cnt].pos[i] = b->pos[i] + vec[i] - axis[i] - n[i] * dd * 0.5;
    cnt++;
  }
  double prjvec1 = -prjvec * 0.5;
  if (dist0 + prjaxis + prjvec1 <= margin) {
    double vec1[3]; v3cross(vec1, vec, axis); v3normalize(vec1); v3scl(vec1, vec1, b->size[0] * sqrt(3.0) / 2);
    dd = dist0 + prjaxis + prjvec1;
    for (int s = -1; s <= 1; s += 2) {
      o[cnt].dist = dd; v3copy(o[cnt].normal, n);
      for (int i = 0; i < 3; i++) o[cnt].pos[i] = b->pos[i] + s * vec1[i] + axis[i] - vec[i] * 0.5 - n[i] * dd * 0.5;
      cnt++;
    }
  }
  return cnt;
}
static int plane_box(const Shape* a, const Shape* b, double margin, RawCon* o) {
  double n[3], t[3]; mat_col(n, a->mat, 2); v3sub(t, b->pos, a->pos);
  double dist = v3dot(t, n);
  int cnt = 0;
  for (int i = 0; i < 8 && cnt < 4; i++) {
    double v[3] = {(i & 1 ? 1 : -1) * b->size[0], (i & 2 ? 1 : -1) * b->size[1], (i & 4 ? 1 : -1) * b->size[2]}, c[3];
    mat_mulvec(c, b->mat, v);
    double ld = v3dot(n, c);
    if (dist + ld > margin) continue;
    o[cnt].dist = dist + ld; v3copy(o[cnt].normal, n);
    for (int k = 0; k < 3; k++) o[cnt].pos[k] = b->pos[k] + c[k] - n[k] * o[cnt].dist * 0.5;
    cnt++;
  }
  return cnt;
}
static int plane_mesh(const Shape* a, const Shape* b, double margin, RawCon* o) {
  double n[3], nl[3], t[3]; mat_col(n, a->mat, 2);
  mat_tmulvec(nl, b->mat, n);
  int best = 0; double bv = 1e300;
  for (int i = 0; i < b->nvert; i++) { double s = v3dot(nl, b->vert + 3 * i); if (s < bv) { bv = s; best = i; } }
  double w[3]; mat_mulvec(w, b->mat, b->vert + 3 * best); v3add(w, w, b->pos);
  v3sub(t, w, a->pos);
  double dist = v3dot(t, n);
  if (dist > margin) return 0;
  o->dist = dist; v3copy(o->normal, n); v3addscl(o->pos, w, n, -0.5 * dist);
  return 1;
}

/* ------------------------------------------------------------------ sphere / capsule pairs */
static int sphere_sphere_raw(const double* c1, double r1, const double* c2, double r2, double margin, RawCon* o) {
  double dif[3]; v3sub(dif, c2, c1);
  double len = v3norm(dif), dist = len - r1 - r2;
  if (dist > margin) return 0;
  if (len < MINVAL) { dif[0] = 1; dif[1] = dif[2] = 0; } else v3scl(dif, dif, 1 / len);
  o->dist = dist; v3copy(o->normal, dif);
  v3addscl(o->pos, c1, dif, r1 + 0.5 * dist);
  return 1;
}
static int sphere_sphere(const Shape* a, const Shape* b, double margin, RawCon* o) {
  return sphere_sphere_raw(a->pos, a->size[0], b->pos, b->size[0], margin, o);
}
static int sphere_capsule(const Shape* a, const Shape* b, double margin, RawCon* o) {
  double ax[3], t[3], c[3]; mat_col(ax, b->mat, 2); v3sub(t, a->pos, b->pos);
  double x = v3dot(t, ax);
  if (x > b->size[1]) x = b->size[1]; if (x < -b->size[1]) x = -b->size[1];
  v3addscl(c, b->pos, ax, x);
  return sphere_sphere_raw(a->pos, a->size[0], c, b->size[0], margin, o);
}
static int capsule_capsule(const Shape* a, const Shape* b, double margin, RawCon* o) {
  double a1[3], a2[3], dif[3]; mat_col(a1, a->mat, 2); mat_col(a2, b->mat, 2); v3sub(dif, a->pos, b->pos);
  double h1 = a->size[1], h2 = b->size[1];
  double ma = v3dot(a1, a1), mb = -v3dot(a1, a2), mc = v3dot(a2, a2), u = -v3dot(a1, dif), v = v3dot(a2, dif);
  double det = ma * mc - mb * mb;
  if (fabs(det) >= 1e-12) {
    double x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
    if (x1 > h1) { x1 = h1; x2 = (v - mb * h1) / mc; } else if (x1 < -h1) { x1 = -h1; x2 = (v + mb * h1) / mc; }
    if (x2 > h2) { x2 = h2; x1 = (u - mb * h2) / ma; if (x1 > h1) x1 = h1; else if (x1 < -h1) x1 = -h1; }
    else if (x2 < -h2) { x2 = -h2; x1 = (u + mb * h2) / ma; if (x1 > h1) x1 = h1; else if (x1 < -h1) x1 = -h1; }
    double p1[3], p2[3]; v3addscl(p1, a->pos, a1, x1); v3addscl(p2, b->pos, a2, x2);
    return sphere_sphere_raw(p1, a->size[0], p2, b->size[0], margin, o);
  }
  /* parallel axes: up to two contacts at the ends of the overlap */
  int cnt = 0;
  double x[2] = {h1, -h1};
  for (int k = 0; k < 2; k++) {
    double p1[3], p2[3], t[3]; v3addscl(p1, a->pos, a1, x[k]); v3sub(t, p1, b->pos);
    double x2 = v3dot(t, a2); if (x2 > h2) x2 = h2; if (x2 < -h2) x2 = -h2;
    v3addscl(p2, b->pos, a2, x2);
    cnt += sphere_sphere_raw(p1, a->size[0], p2, b->size[0], margin, o + cnt);
  }
  return cnt;
}
/* sphere (centre c, radius r) against a box; normal from sphere to box */
static int sphere_box_raw(const double* c, double r, const Shape* b, double margin, RawCon* o) {
  double t[3], cl[3], clamped[3]; v3sub(t, c, b->pos); mat_tmulvec(cl, b->mat, t);
  int inside = 1;
  for (int i = 0; i < 3; i++) {
    clamped[i] = cl[i];
    if (clamped[i] > b->size[i]) { clamped[i] = b->size[i]; inside = 0; }
    else if (clamped[i] < -b->size[i]) { clamped[i] = -b->size[i]; inside = 0; }
  }
  double nl[3], dist, pl[3];
  if (!inside) {
    v3sub(nl, clamped, cl); /* from sphere centre toward the box */
    double len = v3normalize(nl);
    dist = len - r;
    if (dist > margin) return 0;
    v3addscl(pl, cl, nl, r + 0.5 * dist);
  } else {
    /* centre inside: exit through the nearest face */
    int k = 0; double best = 1e300, sgn = 1;
    for (int i = 0; i < 3; i++) {
      double dpos = b->size[i] - cl[i], dneg = b->size[i] + cl[i];
      if (dpos < best) { best = dpos; k = i; sgn = 1; }
      if (dneg < best) { best = dneg; k = i; sgn = -1; }
    }
    v3zero(nl); nl[k] = -sgn;     /* the box lies opposite to the sphere's way out */
    dist = -(best + r);
    v3copy(pl, cl); pl[k] = cl[k] + sgn * 0.5 * (best - r);
  }
  o->dist = dist;
  mat_mulvec(o->normal, b->mat, nl);
  mat_mulvec(o->pos, b->mat, pl); v3add(o->pos, o->pos, b->pos);
  return 1;
}
static int sphere_box(const Shape* a, const Shape* b, double margin, RawCon* o) {
  return sphere_box_raw(a->pos, a->size[0], b, margin, o);
}
/* squared distance from the segment point p0 + t*dir (box frame) to the box, minimised over t in [0,1].
   Piecewise quadratic in t: exact minimisation interval by interval. Returns the minimiser interval. */
static double seg_box_min(const double* p0, const double* dir, const double* size, double* tlo, double* thi) {
  double bp[16]; int nb = 0;
  bp[nb++] = 0; bp[nb++] = 1;
  for (int i = 0; i < 3; i++)
    if (fabs(dir[i]) > 1e-14)
      for (int s = -1; s <= 1; s += 2) { double t = (s * size[i] - p0[i]) / dir[i]; if (t > 0 && t < 1) bp[nb++] = t; }
  for (int i = 1; i < nb; i++) { double x = bp[i]; int j = i - 1; while (j >= 0 && bp[j] > x) { bp[j + 1] = bp[j]; j--; } bp[j + 1] = x; }
  double best = 1e300, blo = 0, bhi = 0;
  for (int k = 0; k + 1 < nb; k++) {
    double t0 = bp[k], t1 = bp[k + 1], tm = 0.5 * (t0 + t1);
    /* f(t) = sum_i (a_i + b_i t)^2 over coordinates outside the slab on this interval */
    double A = 0, B = 0, C = 0;
    for (int i = 0; i < 3; i++) {
      double x = p0[i] + tm * dir[i];
      double off = x > size[i] ? -size[i] : (x < -size[i] ? size[i] : 0);
      if (x > size[i] || x < -size[i]) { double a = p0[i] + off, b = dir[i]; A += b * b; B += 2 * a * b; C += a * a; }
    }
    double tl, th, f;
    double f0 = A * t0 * t0 + B * t0 + C, f1 = A * t1 * t1 + B * t1 + C;
    double ts = A > 0 ? -B / (2 * A) : t0; if (ts < t0) ts = t0; if (ts > t1) ts = t1;
    f = A * ts * ts + B * ts + C;
    if (f0 < 0) f0 = 0; if (f1 < 0) f1 = 0; if (f < 0) f = 0;   /* a squared distance: clear negative round-off */
    if (fmax(f0, f1) - f <= 1e-9 * f + 1e-18) { tl = t0; th = t1; f = fmin(f0, f1) < f ? fmin(f0, f1) : f; } /* flat: segment parallel to the face */
    else { tl = th = ts; }
    if (f < best - (1e-9 * best + 1e-18)) { best = f; blo = tl; bhi = th; }
    else if (fabs(f - best) <= 1e-9 * best + 1e-18 && tl <= bhi + 1e-12) { if (th > bhi) bhi = th; }
  }
  *tlo = blo; *thi = bhi;
  return best;
}
static int capsule_box(const Shape* a, const Shape* b, double margin, RawCon* o) {
  double ax[3], t[3], p0[3], p1[3], l0[3], l1[3], dir[3];
  mat_col(ax, a->mat, 2);
  v3addscl(p0, a->pos, ax, -a->size[1]); v3addscl(p1, a->pos, ax, a->size[1]);
  v3sub(t, p0, b->pos); mat_tmulvec(l0, b->mat, t);
  v3sub(t, p1, b->pos); mat_tmulvec(l1, b->mat, t);
  v3sub(dir, l1, l0);
  double tlo, thi;
  double best = seg_box_min(l0, dir, b->size, &tlo, &thi);
  int cnt = 0;
  double c[3];
  if (best <= 1e-18) {
    /* The capsule axis passes through the box: the distance is zero along [tlo, thi] and the witness direction is
       undefined.  Rule (same on the device): the box face of minimum depth at the midpoint of that interval is the
       contact face for both interval ends; depth is measured at each end. */
    double cm[3], tm = 0.5 * (tlo + thi);
    for (int i = 0; i < 3; i++) cm[i] = l0[i] + tm * dir[i];
    int k = 0; double bd = 1e300, sgn = 1;
    for (int i = 0; i < 3; i++) {
      double dpos = b->size[i] - cm[i], dneg = b->size[i] + cm[i];
      if (dpos < bd) { bd = dpos; k = i; sgn = 1; }
      if (dneg < bd) { bd = dneg; k = i; sgn = -1; }
    }
    int nend = thi - tlo > 1e-9 ? 2 : 1;
    for (int e = 0; e < nend; e++) {
      double te = e ? thi : tlo, ce[3], nl[3] = {0, 0, 0}, pl[3];
      for (int i = 0; i < 3; i++) ce[i] = l0[i] + te * dir[i];
      double depth = b->size[k] - sgn * ce[k]; if (depth < 0) depth = 0;
      nl[k] = -sgn;
      v3copy(pl, ce); pl[k] = ce[k] + sgn * 0.5 * (depth - a->size[0]);
      RawCon* r = o + cnt++;
      r->dist = -(depth + a->size[0]);
      mat_mulvec(r->normal, b->mat, nl);
      mat_mulvec(r->pos, b->mat, pl); v3add(r->pos, r->pos, b->pos);
    }
    return cnt;
  }
  if (thi - tlo > 1e-9) {
    v3addscl(c, p0, ax, 2 * a->size[1] * tlo); cnt += sphere_box_raw(c, a->size[0], b, margin, o + cnt);
    v3addscl(c, p0, ax, 2 * a->size[1] * thi); cnt += sphere_box_raw(c, a->size[0], b, margin, o + cnt);
  } else {
    v3addscl(c, p0, ax, 2 * a->size[1] * tlo); cnt += sphere_box_raw(c, a->size[0], b, margin, o + cnt);
  }
  return cnt;
}

/* ------------------------------------------------------------------ box-box: separating axes + face clipping */
static int clip_poly(double* poly, int n, int axis, double lim, double sgn) {
  /* clip 2D polygon (x,y pairs) against sgn*p[axis] <= lim */
  double out[32]; int m = 0;
  for (int i = 0; i < n; i++) {
    double* p = poly + 2 * i; double* q = poly + 2 * ((i + 1) % n);
    double dp = sgn * p[axis] - lim, dq = sgn * q[axis] - lim;
    if (dp <= 0) { out[2 * m] = p[0]; out[2 * m + 1] = p[1]; m++; }
    if ((dp < 0 && dq > 0) || (dp > 0 && dq < 0)) {
      double s = dp / (dp - dq);
      out[2 * m] = p[0] + s * (q[0] - p[0]); out[2 * m + 1] = p[1] + s * (q[1] - p[1]); m++;
    }
    if (m >= 15) break;
  }
  memcpy(poly, out, sizeof(double) * 2 * (size_t)m);
  return m;
}
static int box_box(const Shape* a, const Shape* b, double margin, RawCon* o) {
  const double* R1 = a->mat; const double* R2 = b->mat;
  double p[3], pp[3]; v3sub(p, b->pos, a->pos); mat_tmulvec(pp, R1, p);
  double R[9], Q[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double c1[3], c2[3]; mat_col(c1, R1, i); mat_col(c2, R2, j);
    R[3 * i + j] = v3dot(c1, c2); Q[3 * i + j] = fabs(R[3 * i + j]);
  }
  const double* A = a->size; const double* B = b->size;
  double s = -1e300; int code = 0, invert = 0; double normalC[3] = {0, 0, 0}; int haveC = 0;
  /* face axes of box 1 */
  for (int i = 0; i < 3; i++) {
    double e = fabs(pp[i]) - (A[i] + B[0] * Q[3 * i] + B[1] * Q[3 * i + 1] + B[2] * Q[3 * i + 2]);
    if (e > margin) return 0;
    if (e > s) { s = e; code = i + 1; invert = pp[i] < 0; haveC = 0; }
  }
  /* face axes of box 2 */
  for (int j = 0; j < 3; j++) {
    double c2[3]; mat_col(c2, R2, j);
    double e1 = v3dot(c2, p);
    double e = fabs(e1) - (A[0] * Q[j] + A[1] * Q[3 + j] + A[2] * Q[6 + j] + B[j]);
    if (e > margin) return 0;
    if (e > s) { s = e; code = j + 4; invert = e1 < 0; haveC = 0; }
  }
  /* edge x edge axes (slightly disfavoured, as in the classic clipping algorithm) */
  const double fudge = 1.05;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    double n[3] = {0, 0, 0};
    n[i1] = -R[3 * i2 + j]; n[i2] = R[3 * i1 + j];
    double l = sqrt(n[i1] * n[i1] + n[i2] * n[i2]);
    if (l < 1e-8) continue;
    double e1 = pp[i2] * R[3 * i1 + j] - pp[i1] * R[3 * i2 + j];
    double e = fabs(e1) - (A[i1] * Q[3 * i2 + j] + A[i2] * Q[3 * i1 + j] + B[j1] * Q[3 * i + j2] + B[j2] * Q[3 * i + j1]);
    e /= l;
    if (e > margin) return 0;
    if ((e < 0 ? e * fudge : e) > s) {
      s = e; code = 7 + 3 * i + j; invert = e1 < 0; haveC = 1;
      normalC[0] = n[0] / l; normalC[1] = n[1] / l; normalC[2] = n[2] / l;
    }
  }
  if (!code) return 0;
  double normal[3];
  if (haveC) mat_mulvec(normal, R1, normalC);
  else if (code <= 3) mat_col(normal, R1, code - 1);
  else mat_col(normal, R2, code - 4);
  if (invert) v3scl(normal, normal, -1);
  double depth = -s; /* >0: penetration */
  if (code > 6) {
    /* edge-edge: closest points between the two edges */
    double pa[3], pb[3]; v3copy(pa, a->pos); v3copy(pb, b->pos);
    for (int j = 0; j < 3; j++) {
      double c1[3], c2[3]; mat_col(c1, R1, j); mat_col(c2, R2, j);
      double sg = v3dot(normal, c1) > 0 ? 1.0 : -1.0; v3addscl(pa, pa, c1, sg * A[j]);
      sg = v3dot(normal, c2) > 0 ? -1.0 : 1.0; v3addscl(pb, pb, c2, sg * B[j]);
    }
    int ia = (code - 7) / 3, ib = (code - 7) % 3;
    double ua[3], ub[3]; mat_col(ua, R1, ia); mat_col(ub, R2, ib);
    double d[3]; v3sub(d, pb, pa);
    double uaub = v3dot(ua, ub), q1 = v3dot(ua, d), q2 = -v3dot(ub, d), dd = 1 - uaub * uaub;
    double alpha = 0, beta = 0;
    if (dd > 1e-4) { alpha = (q1 + uaub * q2) / dd; beta = (uaub * q1 + q2) / dd; }
    v3addscl(pa, pa, ua, alpha); v3addscl(pb, pb, ub, beta);
    o->dist = -depth; v3copy(o->normal, normal);
    for (int k = 0; k < 3; k++) o->pos[k] = 0.5 * (pa[k] + pb[k]);
    return 1;
  }
  /* face contact: reference box = the one owning the axis, incident = the other */
  const double *Ra, *Rb, *pa, *pb, *Sa, *Sb; double nrm[3];
  if (code <= 3) { Ra = R1; Rb = R2; pa = a->pos; pb = b->pos; Sa = A; Sb = B; v3copy(nrm, normal); }
  else { Ra = R2; Rb = R1; pa = b->pos; pb = a->pos; Sa = B; Sb = A; v3scl(nrm, normal, -1); }
  /* incident face: the face of box b most anti-parallel to nrm */
  double nr[3], anr[3]; mat_tmulvec(nr, Rb, nrm);
  for (int k = 0; k < 3; k++) anr[k] = fabs(nr[k]);
  int lanr, a1, a2;
  if (anr[1] > anr[0]) { if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; } }
  else { if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; } }
  double center[3], col[3]; mat_col(col, Rb, lanr);
  double sg = nr[lanr] < 0 ? 1.0 : -1.0;
  for (int k = 0; k < 3; k++) center[k] = pb[k] - pa[k] + sg * Sb[lanr] * col[k];
  /* reference face axes */
  int codeN = (code <= 3 ? code - 1 : code - 4), code1, code2;
  if (codeN == 0) { code1 = 1; code2 = 2; } else if (codeN == 1) { code1 = 0; code2 = 2; } else { code1 = 0; code2 = 1; }
  double r1[3], r2[3], i1v[3], i2v[3];
  mat_col(r1, Ra, code1); mat_col(r2, Ra, code2); mat_col(i1v, Rb, a1); mat_col(i2v, Rb, a2);
  double c1 = v3dot(center, r1), c2 = v3dot(center, r2);
  double m11 = v3dot(r1, i1v), m12 = v3dot(r1, i2v), m21 = v3dot(r2, i1v), m22 = v3dot(r2, i2v);
  double k1 = m11 * Sb[a1], k2 = m21 * Sb[a1], k3 = m12 * Sb[a2], k4 = m22 * Sb[a2];
  double quad[32] = {c1 - k1 - k3, c2 - k2 - k4, c1 - k1 + k3, c2 - k2 + k4, c1 + k1 + k3, c2 + k2 + k4, c1 + k1 - k3, c2 + k2 - k4};
  int n = 4;
  n = clip_poly(quad, n, 0, Sa[code1], 1); if (n) n = clip_poly(quad, n, 0, Sa[code1], -1);
  if (n) n = clip_poly(quad, n, 1, Sa[code2], 1); if (n) n = clip_poly(quad, n, 1, Sa[code2], -1);
  if (n < 1) return 0;
  double det1 = 1.0 / (m11 * m22 - m12 * m21);
  m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
  int cnt = 0;
  for (int j = 0; j < n && cnt < 8; j++) {
    double kk1 = m22 * (quad[2 * j] - c1) - m12 * (quad[2 * j + 1] - c2);
    double kk2 = -m21 * (quad[2 * j] - c1) + m11 * (quad[2 * j + 1] - c2);
    double pt[3];
    for (int k = 0; k < 3; k++) pt[k] = center[k] + kk1 * i1v[k] + kk2 * i2v[k];
    double dep = Sa[codeN] - v3dot(nrm, pt); /* >0: below the reference face */
    if (dep < -margin) continue;
    /* drop duplicates */
    int dup = 0;
    for (int q = 0; q < cnt; q++) {
      double dx[3] = {o[q].pos[0] - (pt[0] + pa[0] + 0.5 * dep * nrm[0]), o[q].pos[1] - (pt[1] + pa[1] + 0.5 * dep * nrm[1]), o[q].pos[2] - (pt[2] + pa[2] + 0.5 * dep * nrm[2])};
      if (v3dot(dx, dx) < 1e-16) dup = 1;
    }
    if (dup) continue;
    o[cnt].dist = -dep; v3copy(o[cnt].normal, normal);
    for (int k = 0; k < 3; k++) o[cnt].pos[k] = pt[k] + pa[k] + 0.5 * dep * nrm[k];
    cnt++;
  }
  return cnt;
}

/* ------------------------------------------------------------------ general convex: GJK + EPA */
typedef struct { double v[3], a[3], b[3]; } SV;

static void support_local(const Shape* s, const double* dl, double* out) {
  switch (s->type) {
    case G_SPHERE: v3zero(out); break; /* core = point, radius added afterwards */
    case G_CAPSULE: out[0] = out[1] = 0; out[2] = dl[2] >= 0 ? s->size[1] : -s->size[1]; break; /* core = segment */
    case G_CYLINDER: {
      double n = sqrt(dl[0] * dl[0] + dl[1] * dl[1]);
      if (n > MINVAL) { out[0] = dl[0] * s->size[0] / n; out[1] = dl[1] * s->size[0] / n; } else out[0] = out[1] = 0;
      out[2] = dl[2] >= 0 ? s->size[1] : -s->size[1];
    } break;
    case G_BOX: for (int i = 0; i < 3; i++) out[i] = dl[i] >= 0 ? s->size[i] : -s->size[i]; break;
    case G_MESH: {
      int best = 0; double bv = -1e300;
      for (int i = 0; i < s->nvert; i++) { double x = v3dot(dl, s->vert + 3 * i); if (x > bv) { bv = x; best = i; } }
      v3copy(out, s->vert + 3 * best);
    } break;
    default: v3zero(out);
  }
}
static double core_radius(const Shape* s) { return (s->type == G_SPHERE || s->type == G_CAPSULE) ? s->size[0] : 0.0; }
static void support(const Shape* A, const Shape* B, const double* dir, SV* o) {
  FL(60 + 6 * ((A->type == G_MESH ? A->nvert : 4) + (B->type == G_MESH ? B->nvert : 4)));
  double dl[3], l[3], nd[3] = {-dir[0], -dir[1], -dir[2]};
  mat_tmulvec(dl, A->mat, dir); support_local(A, dl, l); mat_mulvec(o->a, A->mat, l); v3add(o->a, o->a, A->pos);
  mat_tmulvec(dl, B->mat, nd); support_local(B, dl, l); mat_mulvec(o->b, B->mat, l); v3add(o->b, o->b, B->pos);
  v3sub(o->v, o->a, o->b);
}
/* closest point to the origin on a triangle; barycentric weights out */
static void closest_tri(const double* a, const double* b, const double* c, double* w) {
  double ab[3], ac[3], ap[3]; v3sub(ab, b, a); v3sub(ac, c, a); v3scl(ap, a, -1);
  double d1 = v3dot(ab, ap), d2 = v3dot(ac, ap);
  if (d1 <= 0 && d2 <= 0) { w[0] = 1; w[1] = w[2] = 0; return; }
  double bp[3]; v3scl(bp, b, -1);
  double d3 = v3dot(ab, bp), d4 = v3dot(ac, bp);
  if (d3 >= 0 && d4 <= d3) { w[1] = 1; w[0] = w[2] = 0; return; }
  double vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { double v = d1 / (d1 - d3); w[0] = 1 - v; w[1] = v; w[2] = 0; return; }
  double cp[3]; v3scl(cp, c, -1);
  double d5 = v3dot(ab, cp), d6 = v3dot(ac, cp);
  if (d6 >= 0 && d5 <= d6) { w[2] = 1; w[0] = w[1] = 0; return; }
  double vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { double x = d2 / (d2 - d6); w[0] = 1 - x; w[1] = 0; w[2] = x; return; }
  double va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) { double x = (d4 - d3) / ((d4 - d3) + (d5 - d6)); w[0] = 0; w[1] = 1 - x; w[2] = x; return; }
  double den = 1.0 / (va + vb + vc);
  w[1] = vb * den; w[2] = vc * den; w[0] = 1 - w[1] - w[2];
}
/* reduce simplex to the feature closest to the origin; returns 1 if the origin is enclosed (n==4 only) */
static int closest_simplex(SV* s, int* n, double* v) {
  double w[4] = {0, 0, 0, 0};
  if (*n == 1) { w[0] = 1; }
  else if (*n == 2) {
    double ab[3]; v3sub(ab, s[1].v, s[0].v);
    double t = -v3dot(s[0].v, ab) / fmax(v3dot(ab, ab), 1e-300);
    if (t <= 0) { w[0] = 1; } else if (t >= 1) { w[1] = 1; } else { w[0] = 1 - t; w[1] = t; }
  } else if (*n == 3) {
    closest_tri(s[0].v, s[1].v, s[2].v, w);
  } else {
    /* tetrahedron: test the faces that see the origin */
    static const int F[4][4] = {{0, 1, 2, 3}, {0, 2, 3, 1}, {0, 3, 1, 2}, {1, 3, 2, 0}};
    double best = 1e300; int found = 0; double bw[4] = {0, 0, 0, 0};
    for (int f = 0; f < 4; f++) {
      const double *a = s[F[f][0]].v, *b = s[F[f][1]].v, *c = s[F[f][2]].v, *dv = s[F[f][3]].v;
      double ab[3], ac[3], nrm[3], ad[3]; v3sub(ab, b, a); v3sub(ac, c, a); v3cross(nrm, ab, ac); v3sub(ad, dv, a);
      double sd = v3dot(nrm, ad), so = -v3dot(nrm, a);
      if (fabs(sd) < 1e-300) { so = 1; sd = -1; } /* degenerate: treat the face as visible */
      if ((sd > 0 && so > 0) || (sd < 0 && so < 0)) continue; /* origin on the inner side */
      double tw[3]; closest_tri(a, b, c, tw);
      double q[3]; for (int k = 0; k < 3; k++) q[k] = tw[0] * a[k] + tw[1] * b[k] + tw[2] * c[k];
      double dd = v3dot(q, q);
      if (dd < best) { best = dd; found = 1; memset(bw, 0, sizeof(bw)); bw[F[f][0]] = tw[0]; bw[F[f][1]] = tw[1]; bw[F[f][2]] = tw[2]; }
    }
    if (!found) { v3zero(v); return 1; }
    memcpy(w, bw, sizeof(w));
  }
  SV out[4]; int m = 0; double ww[4];
  v3zero(v);
  for (int i = 0; i < *n; i++) if (w[i] > 0) { out[m] = s[i]; ww[m] = w[i]; for (int k = 0; k < 3; k++) v[k] += w[i] * s[i].v[k]; m++; }
  for (int i = 0; i < m; i++) s[i] = out[i];
  /* stash weights in a scratch slot: recomputed by the caller when needed */
  (void)ww;
  *n = m;
  return 0;
}
static void simplex_weights(const SV* s, int n, double* w) {
  if (n == 1) { w[0] = 1; return; }
  if (n == 2) {
    double ab[3]; v3sub(ab, s[1].v, s[0].v);
    double t = -v3dot(s[0].v, ab) / fmax(v3dot(ab, ab), 1e-300);
    if (t < 0) t = 0; if (t > 1) t = 1; w[0] = 1 - t; w[1] = t; return;
  }
  closest_tri(s[0].v, s[1].v, s[2].v, w);
}

#define EPA_MAXV 64
#define EPA_MAXF 320
#define GJK_TOL 1e-8       /* absolute gap between the GJK upper and lower distance bounds */
#define EPA_ITERS 50
#define EPA_TOL 1e-6
typedef struct { int v[3]; double n[3], d; int alive; } EFace;

static int epa_add_face(EFace* F, int* nf, const SV* V, int a, int b, int c) {
  if (*nf >= EPA_MAXF) return -1;
  EFace* f = F + (*nf);
  f->v[0] = a; f->v[1] = b; f->v[2] = c; f->alive = 1;
  double ab[3], ac[3]; v3sub(ab, V[b].v, V[a].v); v3sub(ac, V[c].v, V[a].v); v3cross(f->n, ab, ac);
  double l = v3norm(f->n);
  if (l < 1e-300) { f->d = 0; f->n[0] = 1; f->n[1] = f->n[2] = 0; }
  else { v3scl(f->n, f->n, 1 / l); f->d = v3dot(f->n, V[a].v); }
  if (f->d < 0) { /* keep normals pointing away from the origin */
    int t = f->v[1]; f->v[1] = f->v[2]; f->v[2] = t; v3scl(f->n, f->n, -1); f->d = -f->d;
  }
  (*nf)++;
  return 0;
}

/* returns 1 with (dist<=0 penetration) filled, 0 on failure */
static int epa(const Shape* A, const Shape* B, SV* s, int n, RawCon* o, double* wa, double* wb) {
  SV V[EPA_MAXV]; EFace F[EPA_MAXF]; int nv = 0, nf = 0;
  /* grow the simplex to a tetrahedron */
  static const double dirs[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  if (n == 1) { for (int k = 0; k < 6 && n < 2; k++) { SV w; support(A, B, dirs[k], &w); double dd[3]; v3sub(dd, w.v, s[0].v); if (v3dot(dd, dd) > 1e-20) s[n++] = w; } }
  if (n == 2) {
    double ab[3]; v3sub(ab, s[1].v, s[0].v);
    for (int k = 0; k < 6 && n < 3; k++) {
      double dir[3]; v3cross(dir, ab, dirs[k]);
      if (v3dot(dir, dir) < 1e-12 * v3dot(ab, ab)) continue;
      for (int sg = 0; sg < 2 && n < 3; sg++) {
        SV w; support(A, B, dir, &w);
        double aw[3], cr[3]; v3sub(aw, w.v, s[0].v); v3cross(cr, ab, aw);
        if (v3dot(cr, cr) > 1e-20) s[n++] = w;
        v3scl(dir, dir, -1);
      }
    }
  }
  if (n == 3) {
    double ab[3], ac[3], nn[3]; v3sub(ab, s[1].v, s[0].v); v3sub(ac, s[2].v, s[0].v); v3cross(nn, ab, ac);
    for (int sg = 0; sg < 2 && n < 4; sg++) {
      SV w; support(A, B, nn, &w);
      double aw[3]; v3sub(aw, w.v, s[0].v);
      if (fabs(v3dot(aw, nn)) > 1e-14 * sqrt(v3dot(nn, nn))) s[n++] = w;
      v3scl(nn, nn, -1);
    }
  }
  if (n < 4) return 0;
  for (int i = 0; i < 4; i++) V[nv++] = s[i];
  epa_add_face(F, &nf, V, 0, 1, 2); epa_add_face(F, &nf, V, 0, 2, 3); epa_add_face(F, &nf, V, 0, 3, 1); epa_add_face(F, &nf, V, 1, 3, 2);
  int bestf = -1;
  /* expansion loop; iteration limit and tolerance follow MuJoCo's mjOption defaults for its convex collision pipeline
     (ccd_iterations = 50, ccd_tolerance = 1e-6) [3P] */
  for (int it = 0; it < EPA_ITERS; it++) {
    bestf = -1; double bd = 1e300;
    for (int f = 0; f < nf; f++) if (F[f].alive && F[f].d < bd) { bd = F[f].d; bestf = f; }
    if (bestf < 0) return 0;
    SV w; support(A, B, F[bestf].n, &w);
    double dw = v3dot(w.v, F[bestf].n);
    if (dw - bd < EPA_TOL || nv >= EPA_MAXV) break;
    /* faces visible from w, in index order */
    int vis[EPA_MAXF], nvis = 0;
    for (int f = 0; f < nf; f++) {
      if (!F[f].alive) continue;
      double t[3]; v3sub(t, w.v, V[F[f].v[0]].v);
      if (v3dot(F[f].n, t) > 1e-14) vis[nvis++] = f;
    }
    /* horizon = edges of visible faces whose reverse is not an edge of a visible face; order: (face, edge) ascending */
    int edges[EPA_MAXF * 3][2]; int ne = 0;
    for (int x = 0; x < nvis; x++)
      for (int e = 0; e < 3; e++) {
        int ea = F[vis[x]].v[e], eb = F[vis[x]].v[(e + 1) % 3], shared = 0;
        for (int y = 0; y < nvis && !shared; y++)
          for (int e2 = 0; e2 < 3; e2++) if (F[vis[y]].v[e2] == eb && F[vis[y]].v[(e2 + 1) % 3] == ea) { shared = 1; break; }
        if (!shared) { edges[ne][0] = ea; edges[ne][1] = eb; ne++; }
      }
    for (int x = 0; x < nvis; x++) F[vis[x]].alive = 0;
    if (ne == 0 || nf + ne > EPA_MAXF) break;
    int wi = nv; V[nv++] = w;
    for (int q = 0; q < ne; q++) epa_add_face(F, &nf, V, edges[q][0], edges[q][1], wi);
  }
  if (bestf < 0) return 0;
  /* witness points from the closest face */
  const EFace* f = F + bestf;
  double w3[3]; closest_tri(V[f->v[0]].v, V[f->v[1]].v, V[f->v[2]].v, w3);
  v3zero(wa); v3zero(wb);
  for (int k = 0; k < 3; k++) { v3addscl(wa, wa, V[f->v[k]].a, w3[k]); v3addscl(wb, wb, V[f->v[k]].b, w3[k]); }
  /* origin inside A-B; the closest boundary point is n*d, so A must retreat along -n: B lies along +n from A */
  o->dist = -f->d;
  v3copy(o->normal, f->n);
  return 1;
}


/* Face-on contact between a cylinder and a box face (cap-on-face or generator-line-on-face): every point of the
   contact patch is a valid EPA witness, so the contact *point* is implementation-defined in the reference physics.
   We take the pressure centroid of the patch under a linear penetration profile, clipped to the overlap with the
   box face (continuous in the pose; removes rim-to-rim chatter of resting cylinders).  Other configurations are
   left as EPA/GJK produced them. */
static void refine_cyl_box(const Shape* A, const Shape* B, RawCon* o) {
  if (A->type != G_CYLINDER || B->type != G_BOX) return;
  double nb[3]; mat_tmulvec(nb, B->mat, o->normal);
  int k = 0; for (int q = 1; q < 3; q++) if (fabs(nb[q]) > fabs(nb[k])) k = q;
  if (fabs(nb[k]) < 0.99999) return;                       /* the contact normal must be a face normal of the box */
  const double sgn = nb[k] > 0 ? 1 : -1;
  const int i = (k + 1) % 3, j = (k + 2) % 3;
  const double r = A->size[0], h = A->size[1];
  double ax[3], a[3], t[3], cB[3];
  mat_col(ax, A->mat, 2); mat_tmulvec(a, B->mat, ax);
  v3sub(t, A->pos, B->pos); mat_tmulvec(cB, B->mat, t);
  const double ak = a[k] * sgn;
  double P[3], pen;
  if (fabs(ak) > 0.9) {
    /* ---- cap on face */
    const double cs = ak > 0 ? 1 : -1;
    double C[3], acap[3], g[3];
    for (int q = 0; q < 3; q++) { acap[q] = cs * a[q]; C[q] = cB[q] + cs * h * a[q]; }
    const double pen0 = sgn * C[k] + B->size[k];
    for (int q = 0; q < 3; q++) g[q] = -(sgn * acap[k]) * acap[q];
    g[k] += sgn;
    const double st = v3norm(g);                                   /* sin(tilt) */
    double sbar = 0;
    if (st > 1e-9) sbar = pen0 > 0 ? fmin(r, r * r * st / (4 * pen0)) : r;
    for (int q = 0; q < 3; q++) P[q] = C[q] + (st > 1e-9 ? g[q] / st * sbar : 0);
    const int axs[2] = {i, j};
    for (int q = 0; q < 2; q++) {
      const int u = axs[q];
      const double lo = fmax(C[u] - r, -B->size[u]), hi = fmin(C[u] + r, B->size[u]);
      if (lo > hi) return;
      P[u] = fmin(fmax(P[u], lo), hi);
    }
    P[k] = C[k] - (acap[i] * (P[i] - C[i]) + acap[j] * (P[j] - C[j])) / acap[k];
    pen = sgn * P[k] + B->size[k];
  } else if (fabs(ak) < 0.1) {
    /* ---- generator line on face */
    double d[3], L0[3];
    for (int q = 0; q < 3; q++) d[q] = -(sgn * a[k]) * a[q];
    d[k] += sgn;
    v3normalize(d);
    for (int q = 0; q < 3; q++) L0[q] = cB[q] + r * d[q];
    double t0 = -h, t1 = h;
    const int axs[2] = {i, j};
    for (int q = 0; q < 2; q++) {
      const int u = axs[q];
      if (fabs(a[u]) > 1e-9) {
        const double ta = (-B->size[u] - L0[u]) / a[u], tb = (B->size[u] - L0[u]) / a[u];
        t0 = fmax(t0, fmin(ta, tb)); t1 = fmin(t1, fmax(ta, tb));
      } else if (fabs(L0[u]) > B->size[u]) return;
    }
    if (t0 > t1) return;
    const double tm = 0.5 * (t0 + t1), Lh = 0.5 * (t1 - t0), sl = sgn * a[k];
    const double penm = sgn * (L0[k] + tm * a[k]) + B->size[k];
    double off;
    if (penm > 0) off = fmin(fmax(sl * Lh * Lh / (3 * penm), -Lh), Lh);
    else off = sl > 0 ? Lh : (sl < 0 ? -Lh : 0);
    const double ts = tm + off;
    for (int q = 0; q < 3; q++) P[q] = L0[q] + ts * a[q];
    pen = sgn * P[k] + B->size[k];
  } else return;
  o->dist = -pen;
  P[k] -= sgn * 0.5 * pen;
  mat_mulvec(o->pos, B->mat, P); v3add(o->pos, o->pos, B->pos);
}

/* Cylinder vs box with the cylinder axis parallel to a box axis (buttons in housings, handles, a puck at rest): the
   problem separates into an interval overlap along the axis and disc-vs-rectangle in the plane across it, so distance,
   penetration depth and normal are exact and cheap.  Returns -1 when the axes are not parallel (GJK/EPA handles it). */
static int cyl_box_aligned(const Shape* A, const Shape* B, double margin, RawCon* o) {
  double ax[3], a[3], t[3], c[3];
  mat_col(ax, A->mat, 2); mat_tmulvec(a, B->mat, ax);
  int k = 0; for (int q = 1; q < 3; q++) if (fabs(a[q]) > fabs(a[k])) k = q;
  if (fabs(a[k]) < 1 - 1e-6) return -1;      /* within ~1.4 mrad: MJCF quaternions like "0.7074 0.7068 0 0" are "parallel" */
  const int i = (k + 1) % 3, j = (k + 2) % 3;
  v3sub(t, A->pos, B->pos); mat_tmulvec(c, B->mat, t);
  const double r = A->size[0], h = A->size[1]; const double* s = B->size;
  const double cz = c[k], sz = cz >= 0 ? 1 : -1;
  const double ga = fabs(cz) - (h + s[k]);                       /* axial gap (negative: overlap) */
  const double p[2] = {c[i], c[j]};
  const double q[2] = {fmin(fmax(p[0], -s[i]), s[i]), fmin(fmax(p[1], -s[j]), s[j])};
  double n2[2], gr;                                              /* in-plane direction box -> cylinder, radial gap */
  if (q[0] != p[0] || q[1] != p[1]) {
    const double dv[2] = {p[0] - q[0], p[1] - q[1]}, dl = sqrt(dv[0] * dv[0] + dv[1] * dv[1]);
    gr = dl - r; n2[0] = dv[0] / dl; n2[1] = dv[1] / dl;
  } else {
    const double ei = s[i] - fabs(p[0]), ej = s[j] - fabs(p[1]);
    if (ei <= ej) { gr = -ei - r; n2[0] = p[0] >= 0 ? 1 : -1; n2[1] = 0; }
    else { gr = -ej - r; n2[0] = 0; n2[1] = p[1] >= 0 ? 1 : -1; }
  }
  double nB[3] = {0, 0, 0}, P[3], dist;
  if (ga > 0 && gr > 0) {                                        /* rim against a box edge */
    dist = sqrt(ga * ga + gr * gr);
    nB[i] = gr * n2[0] / dist; nB[j] = gr * n2[1] / dist; nB[k] = ga * sz / dist;
    P[i] = 0.5 * (p[0] - r * n2[0] + q[0]); P[j] = 0.5 * (p[1] - r * n2[1] + q[1]); P[k] = 0.5 * (cz - sz * h + sz * s[k]);
  } else if (ga > gr) {                                          /* cap against a box face */
    dist = ga; nB[k] = sz;
    P[i] = q[0]; P[j] = q[1]; P[k] = sz * s[k] + 0.5 * ga * sz;
  } else {                                                       /* side against a box face or edge */
    dist = gr; nB[i] = n2[0]; nB[j] = n2[1];
    const double z0 = fmax(cz - h, -s[k]), z1 = fmin(cz + h, s[k]);
    P[i] = p[0] - (r + 0.5 * gr) * n2[0]; P[j] = p[1] - (r + 0.5 * gr) * n2[1]; P[k] = 0.5 * (z0 + z1);
  }
  if (dist > margin + 1e-4) return 0;                            /* slack: the refinement below makes the final call */
  o->dist = dist;
  v3scl(nB, nB, -1);                                             /* normal points from the cylinder (geom1) to the box */
  mat_mulvec(o->normal, B->mat, nB);
  mat_mulvec(o->pos, B->mat, P); v3add(o->pos, o->pos, B->pos);
  refine_cyl_box(A, B, o);
  return o->dist <= margin;
}
/* Two cylinders with parallel axes (the faucet's stacked discs): same separation of variables. */
static int cyl_cyl_parallel(const Shape* A, const Shape* B, double margin, RawCon* o) {
  double a1[3], a2[3], c[3], cr[3], u[3];
  mat_col(a1, A->mat, 2); mat_col(a2, B->mat, 2);
  if (fabs(v3dot(a1, a2)) < 1 - 1e-6) return -1;
  v3sub(c, B->pos, A->pos);
  const double cz = v3dot(c, a1), sz = cz >= 0 ? 1 : -1;
  v3addscl(cr, c, a1, -cz);
  const double rho = v3norm(cr);
  if (rho > 1e-12) v3scl(u, cr, 1 / rho); else mat_col(u, A->mat, 0);
  const double r1 = A->size[0], h1 = A->size[1], r2 = B->size[0], h2 = B->size[1];
  const double ga = fabs(cz) - (h1 + h2), gr = rho - (r1 + r2);
  double dist, n[3], P[3];
  if (ga > 0 && gr > 0) {
    dist = sqrt(ga * ga + gr * gr);
    for (int q = 0; q < 3; q++) {
      n[q] = (gr * u[q] + ga * sz * a1[q]) / dist;
      const double pa = A->pos[q] + r1 * u[q] + sz * h1 * a1[q], pb = B->pos[q] - r2 * u[q] - sz * h2 * a1[q];
      P[q] = 0.5 * (pa + pb);
    }
  } else if (ga > gr) {
    dist = ga;
    const double t0 = fmax(-r1, rho - r2), t1 = fmin(r1, rho + r2), tm = 0.5 * (t0 + t1);
    for (int q = 0; q < 3; q++) { n[q] = sz * a1[q]; P[q] = A->pos[q] + tm * u[q] + (sz * h1 + 0.5 * ga * sz) * a1[q]; }
  } else {
    dist = gr;
    const double z0 = fmax(-h1, cz - h2), z1 = fmin(h1, cz + h2), zm = 0.5 * (z0 + z1);
    for (int q = 0; q < 3; q++) { n[q] = u[q]; P[q] = A->pos[q] + (r1 + 0.5 * gr) * u[q] + zm * a1[q]; }
  }
  if (dist > margin) return 0;
  o->dist = dist; v3copy(o->normal, n); v3copy(o->pos, P);
  return 1;
}

static int convex_convex(const Shape* A, const Shape* B, double margin, RawCon* o) {
  if (A->type == G_CYLINDER && B->type == G_BOX) { int r = cyl_box_aligned(A, B, margin, o); if (r >= 0) return r; }
  if (A->type == G_CYLINDER && B->type == G_CYLINDER) { int r = cyl_cyl_parallel(A, B, margin, o); if (r >= 0) return r; }
  SV s[4]; int n = 0;
  double v[3]; v3sub(v, A->pos, B->pos);
  if (v3dot(v, v) < 1e-24) { v[0] = 1; v[1] = v[2] = 0; }
  double ra = core_radius(A), rb = core_radius(B);
  { double nd[3] = {-v[0], -v[1], -v[2]}; support(A, B, nd, &s[0]); n = 1; v3copy(v, s[0].v); }
  int enclosed = 0;
  for (int it = 0; it < 64; it++) {
    double vv = v3dot(v, v);
    if (vv < 1e-24) { enclosed = 1; break; }
    double nd[3] = {-v[0], -v[1], -v[2]};
    SV w; support(A, B, nd, &w);
    double vw = v3dot(v, w.v);
    if (vv - vw <= 1e-12 * vv || vv - vw <= GJK_TOL * sqrt(vv)) break; /* |v| is within GJK_TOL of the lower bound v.w/|v|: v is the closest point */
    if (vw > 0 && vw / sqrt(vv) - ra - rb > margin + 1e-4) return 0; /* separating axis */
    int dup = 0;
    for (int i = 0; i < n; i++) { double t[3]; v3sub(t, s[i].v, w.v); if (v3dot(t, t) < 1e-24) dup = 1; }
    if (dup) break;
    s[n++] = w;
    FL(150);
    if (closest_simplex(s, &n, v)) { enclosed = 1; break; }
  }
  double wa[3], wb[3];
  if (!enclosed) {
    double w[4]; simplex_weights(s, n, w);
    v3zero(wa); v3zero(wb);
    for (int i = 0; i < n; i++) { v3addscl(wa, wa, s[i].a, w[i]); v3addscl(wb, wb, s[i].b, w[i]); }
    double dvec[3]; v3sub(dvec, wb, wa);
    double dcore = v3norm(dvec);
    if (dcore > 1e-10) {
      double dist = dcore - ra - rb;
      if (dist > margin + 1e-4) return 0;      /* slack: the analytic refinement below makes the final call */
      v3scl(o->normal, dvec, 1 / dcore);
      o->dist = dist;
      double sa[3], sb[3]; v3addscl(sa, wa, o->normal, ra); v3addscl(sb, wb, o->normal, -rb);
      for (int k = 0; k < 3; k++) o->pos[k] = 0.5 * (sa[k] + sb[k]);
      refine_cyl_box(A, B, o);
      return o->dist <= margin;
    }
    enclosed = 1;
  }
  RawCon r;
  if (!epa(A, B, s, n, &r, wa, wb)) return 0;
  double dist = r.dist - ra - rb;
  if (dist > margin) return 0;
  o->dist = dist; v3copy(o->normal, r.normal);
  double sa[3], sb[3]; v3addscl(sa, wa, o->normal, ra); v3addscl(sb, wb, o->normal, -rb);
  for (int k = 0; k < 3; k++) o->pos[k] = 0.5 * (sa[k] + sb[k]);
  refine_cyl_box(A, B, o);
  return o->dist <= margin;
}

/* ------------------------------------------------------------------ driver */
static void make_frame(double* fr) {
  /* fr[0..2] = normal given; complete the tangents  [3P mju_makeFrame] */
  v3normalize(fr);
  double* y = fr + 3; double* z = fr + 6;
  v3zero(y);
  if (fr[1] < 0.5 && fr[1] > -0.5) y[1] = 1; else y[2] = 1;
  double dp = v3dot(fr, y);
  v3addscl(y, y, fr, -dp);
  v3normalize(y);
  v3cross(z, fr, y);
}

static int narrowphase(const Shape* a, const Shape* b, double margin, RawCon* o) {
  int t1 = a->type, t2 = b->type;
  if (t1 == G_PLANE) {
    switch (t2) {
      case G_SPHERE: return plane_sphere(a, b, margin, o);
      case G_CAPSULE: return plane_capsule(a, b, margin, o);
      case G_CYLINDER: return plane_cylinder(a, b, margin, o);
      case G_BOX: return plane_box(a, b, margin, o);
      case G_MESH: return plane_mesh(a, b, margin, o);
      default: return 0;
    }
  }
  if (t1 == G_SPHERE && t2 == G_SPHERE) return sphere_sphere(a, b, margin, o);
  if (t1 == G_SPHERE && t2 == G_CAPSULE) return sphere_capsule(a, b, margin, o);
  if (t1 == G_SPHERE && t2 == G_BOX) return sphere_box(a, b, margin, o);
  if (t1 == G_CAPSULE && t2 == G_CAPSULE) return capsule_capsule(a, b, margin, o);
  if (t1 == G_CAPSULE && t2 == G_BOX) return capsule_box(a, b, margin, o);
  if (t1 == G_BOX && t2 == G_BOX) return box_box(a, b, margin, o);
  return convex_convex(a, b, margin, o);
}

/* test hook: narrowphase of one pair given explicitly (tests/test_devcollide.py fuzzes the device code against it) */
int om_narrowphase_pair(int t1, const double* pos1, const double* mat1, const double* size1, const double* vert1, int nv1,
                        int t2, const double* pos2, const double* mat2, const double* size2, const double* vert2, int nv2,
                        double margin, double* out) {
  Shape a = {t1, pos1, mat1, size1, vert1, nv1}, b = {t2, pos2, mat2, size2, vert2, nv2};
  RawCon rc[16];
  int cnt = narrowphase(&a, &b, margin, rc);
  for (int i = 0; i < cnt; i++) { out[7 * i] = rc[i].dist; for (int k = 0; k < 3; k++) { out[7 * i + 1 + k] = rc[i].pos[k]; out[7 * i + 4 + k] = rc[i].normal[k]; } }
  return cnt;
}

void om_collide(const OModel* m, OData* d) {
  d->ncon = 0;
  RawCon raw[16];
  for (int p = 0; p < m->npair; p++) {
    int g1 = m->pair_g1[p], g2 = m->pair_g2[p];
    double margin = fmax(m->geom_margin[g1], m->geom_margin[g2]);
    double gap = fmax(m->geom_gap[g1], m->geom_gap[g2]);
    Shape a, b; get_shape(m, d, g1, &a); get_shape(m, d, g2, &b);
    FL(12);
    /* bounding-sphere cull (a plane has no bound) */
    if (a.type == G_PLANE) {
      double n[3], t[3]; mat_col(n, a.mat, 2); v3sub(t, b.pos, a.pos);
      if (v3dot(t, n) > m->geom_rbound[g2] + margin) continue;
    } else {
      double t[3]; v3sub(t, b.pos, a.pos);
      double bound = m->geom_rbound[g1] + m->geom_rbound[g2] + margin;
      if (v3dot(t, t) > bound * bound) continue;
    }
    FL(a.type == G_BOX && b.type == G_BOX ? 900 : 120);      /* analytic narrowphase (box-box SAT + clipping is the large one); GJK/EPA add their own below */
    int n = narrowphase(&a, &b, margin, raw);
    for (int k = 0; k < n && d->ncon < OM_MAXCON; k++) {
      OContact* c = d->contact + d->ncon++;
      memset(c, 0, sizeof(*c));
      c->dist = raw[k].dist; v3copy(c->pos, raw[k].pos); v3copy(c->frame, raw[k].normal);
      make_frame(c->frame);
      c->geom1 = g1; c->geom2 = g2; c->includemargin = margin - gap; c->efc_address = -1;
      /* parameter mixing  [3P mj_contactParam], equal priorities in every Meta-World model */
      int p1 = m->geom_priority[g1], p2 = m->geom_priority[g2];
      const double* f1 = m->geom_friction + 3 * g1; const double* f2 = m->geom_friction + 3 * g2;
      double fr[3];
      if (p1 != p2) {
        int gp = p1 > p2 ? g1 : g2;
        c->dim = m->geom_condim[gp];
        memcpy(c->solref, m->geom_solref + 2 * gp, 2 * sizeof(double));
        memcpy(c->solimp, m->geom_solimp + 5 * gp, 5 * sizeof(double));
        memcpy(fr, m->geom_friction + 3 * gp, 3 * sizeof(double));
      } else {
        c->dim = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
        double s1 = m->geom_solmix[g1], s2 = m->geom_solmix[g2], mix;
        if (s1 >= MINVAL && s2 >= MINVAL) mix = s1 / (s1 + s2);
        else if (s1 < MINVAL && s2 < MINVAL) mix = 0.5;
        else mix = s1 < MINVAL ? 0.0 : 1.0;
        const double* r1 = m->geom_solref + 2 * g1; const double* r2 = m->geom_solref + 2 * g2;
        if (r1[0] > 0 && r2[0] > 0) for (int i = 0; i < 2; i++) c->solref[i] = mix * r1[i] + (1 - mix) * r2[i];
        else for (int i = 0; i < 2; i++) c->solref[i] = fmin(r1[i], r2[i]);
        for (int i = 0; i < 5; i++) c->solimp[i] = mix * m->geom_solimp[5 * g1 + i] + (1 - mix) * m->geom_solimp[5 * g2 + i];
        for (int i = 0; i < 3; i++) fr[i] = fmax(f1[i], f2[i]);
      }
      c->friction[0] = c->friction[1] = fr[0]; c->friction[2] = fr[1]; c->friction[3] = c->friction[4] = fr[2];
      c->mu = fr[0];
    }
  }
}
