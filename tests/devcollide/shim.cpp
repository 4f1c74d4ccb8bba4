// Host emulation of the device narrowphase (metaworld_b200/csrc/mw_collide.cuh) for CPU-side parity fuzzing against
// the oracle: the CUDA source is compiled by g++ with the warp collapsed to ONE lane (MW_WARP = 1, shuffles = identity),
// so every lane-strided loop runs to completion on "lane 0".  Test infrastructure only.
#define MW_HOST_EMU
#define MW_WARP 1
#include <cmath>
#include <cstring>
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int) { return v; }
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline void __syncwarp() {}
template <class T> static inline T __ldg(const T* p) { return *p; }
struct float4 { float x, y, z, w; };
#include "../../metaworld_b200/csrc/mw_collide.cuh"

#include <vector>
static std::vector<float4> g_pack[2];
static void load(DShape* s, int type, const double* pos, const double* mat, const double* size, const float* vert3, int nvert, int slot) {
  g_pack[slot].resize(nvert > 0 ? nvert : 1);
  for (int i = 0; i < nvert; i++) g_pack[slot][i] = float4{vert3[3 * i], vert3[3 * i + 1], vert3[3 * i + 2], 0.f};
  const float4* vert = g_pack[slot].data();
  s->type = type;
  s->pos = pos; s->mat = mat;                                        // DShape points at the pose (the kernel: shared memory)
  for (int i = 0; i < 3; i++) s->size[i] = (float)size[i];           // sizes are float32 in MwModel
  s->vert = vert; s->nvert = nvert;
}

extern "C" int dev_pair(int t1, const double* pos1, const double* mat1, const double* size1, const float* vert1, int nv1,
                        int t2, const double* pos2, const double* mat2, const double* size2, const float* vert2, int nv2,
                        double margin, double* out /* [8][7]: dist, pos, normal */) {
  DShape a, b; load(&a, t1, pos1, mat1, size1, vert1, nv1, 0); load(&b, t2, pos2, mat2, size2, vert2, nv2, 1);
  RawCon rc[8]; int cnt = 0;
  static EpaSm E; static EpaWs W;
  if (pair_is_analytic(t1, t2)) cnt = narrow_analytic(a, b, margin, rc);
  else {
    int r = -1;
    if (t1 == G_CYLINDER && t2 == G_BOX) r = cyl_box_aligned(a, b, margin, rc);
    else if (t1 == G_CYLINDER && t2 == G_CYLINDER) r = cyl_cyl_parallel(a, b, margin, rc);
    if (r >= 0) cnt = r;
    else if (t1 == G_PLANE) {   // plane - mesh (as in mw_collide)
      creal n[3], nd[3], sp[3], t[3]; mat_col(n, a.mat, 2); v3scl(nd, n, -1);
      support_shape(b, nd, sp, 0);
      v3sub(t, sp, a.pos);
      rc[0].dist = v3dot(t, n); v3copy(rc[0].normal, n); v3addscl(rc[0].pos, sp, n, -(creal)0.5 * rc[0].dist);
      cnt = rc[0].dist <= margin;
    } else cnt = convex_pair(a, b, margin, rc, &E, &W, 0, nullptr);
  }
  for (int i = 0; i < cnt; i++) { out[7 * i] = rc[i].dist; for (int k = 0; k < 3; k++) { out[7 * i + 1 + k] = rc[i].pos[k]; out[7 * i + 4 + k] = rc[i].normal[k]; } }
  return cnt;
}
