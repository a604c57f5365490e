// TEST HARNESS (not product): runs the product's per-lane device functions
// (omni3d_b200/csrc/box3d_geom.cuh, compiled for the host) in a serial driver that mirrors the
// warp kernel's stage order, so CPU-only CI can compare them bit-for-bit with the oracle.
#include <vector>
#include <cstring>
#include "../../omni3d_b200/csrc/box3d_geom.cuh"
using namespace c3d;

static V3 ld3(const float* p) { return mk(p[0], p[1], p[2]); }

static int clip_side(const float* recT, const float* recP, std::vector<Tri>& out) {
  std::vector<Tri> cur, nxt;
  for (int t = 0; t < 12; ++t) {
    Tri tr; tr.a = ld3(recT + 3 * tri_vert(t, 0)); tr.b = ld3(recT + 3 * tri_vert(t, 1));
    tr.c = ld3(recT + 3 * tri_vert(t, 2));
    cur.push_back(tr);
  }
  for (int p = 0; p < 6; ++p) {
    V3 pc = ld3(recP + 24 + 6 * p), n = ld3(recP + 27 + 6 * p);
    V3 q[4];
    for (int k = 0; k < 4; ++k) q[k] = ld3(recP + 3 * plane_vert(p, k));
    nxt.clear();
    for (auto& t : cur) {
      Tri o0, o1;
      int k = clip_tri(t, pc, n, q, &o0, &o1);
      if (k >= 1) nxt.push_back(o0);
      if (k == 2) nxt.push_back(o1);
    }
    cur.swap(nxt);
  }
  out = cur;
  return (int)cur.size();
}

extern "C" void harness_iou(const float* b1, int N, const float* b2, int M, float* vol, float* iou,
                            int32_t* nfaces) {
  std::vector<float> r1(64 * (size_t)N), r2(64 * (size_t)M);
  for (int i = 0; i < N; ++i) build_box_record(b1 + 24 * i, &r1[64 * i], nullptr);
  for (int j = 0; j < M; ++j) build_box_record(b2 + 24 * j, &r2[64 * j], nullptr);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < M; ++j) {
      const float* A = &r1[64 * i]; const float* B = &r2[64 * j];
      std::vector<Tri> i1, i2;
      int n1 = clip_side(A, B, i1);
      int n2 = clip_side(B, A, i2);
      std::vector<V3> nr1(n1), nr2(n2); std::vector<float> ar1(n1);
      for (int a = 0; a < n1; ++a) { nr1[a] = tri_normal(i1[a]); ar1[a] = tri_area(i1[a]); }
      for (int b = 0; b < n2; ++b) nr2[b] = tri_normal(i2[b]);
      std::vector<Tri> fin = i1;
      for (int b = 0; b < n2; ++b) {
        bool keep = true;
        for (int a = 0; a < n1; ++a)
          if (ar1[a] > aEps && coplanar_tri_tri(i1[a], nr1[a], i2[b], nr2[b])) keep = false;
        if (keep) fin.push_back(i2[b]);
      }
      float v = 0.f, u = 0.f;
      int nf = (int)fin.size();
      if (nf > 0) {
        float x = 0, y = 0, z = 0;
        for (auto& t : fin) {
          x += (t.a.x + t.b.x + t.c.x) / 3.0f; y += (t.a.y + t.b.y + t.c.y) / 3.0f;
          z += (t.a.z + t.b.z + t.c.z) / 3.0f;
        }
        V3 c = mk(x / nf, y / nf, z / nf);
        for (auto& t : fin) v = v + tet_volume(t, c);
        u = v / (A[63] + B[63] - v);
      }
      size_t k = (size_t)i * M + j;
      vol[k] = v; iou[k] = u; nfaces[k] = nf;
    }
}
extern "C" void harness_check(const float* b, int N, float e1, float e2, int32_t* flags) {
  for (int i = 0; i < N; ++i) flags[i] = check_box(b + 24 * i, e1, e2);
}
