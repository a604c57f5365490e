/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the
 * product path (omni3d_b200/); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * CPU restatement (plain C, fp32, serial double loop) of the arithmetic behind
 *   cubercnn/evaluation/omni3d_evaluation.py:106-166  box3d_overlap
 *   cubercnn/evaluation/omni3d_evaluation.py:65-86    _check_coplanar
 *   cubercnn/evaluation/omni3d_evaluation.py:89-104   _check_nonzero
 * whose heavy lifting is the un-vendored third-party op
 *   pytorch3d._C.iou_box3d  (omni3d_evaluation.py:37,155)
 * PyTorch3D is NOT under /root/reference and not installed here; the reference
 * does not pin a version (README.md:60 -> conda "pytorch3d", >= v0.5 because
 * iou_box3d first shipped there).  The algorithm below restates PyTorch3D's
 * published CPU algorithm (pytorch3d/csrc/iou_box3d/iou_utils.h +
 * iou_box3d_cpu.cpp, main-branch form with the "best-conditioned edge pair"
 * normals): triangle-vs-plane Sutherland-Hodgman clipping both ways, removal
 * of box2-side triangles coplanar with a box1-side triangle, volume by signed
 * tetrahedra about the polyhedron centre.
 *
 * PARITY PIN: the reference holds no tests / golden vectors for this path
 *  -> "parity unpinned" by the reference itself.  This
 * file is pinned instead against (i) closed-form IoU answers and (ii) an
 * independent fp64 half-space-intersection oracle (scipy), see
 * tests/test_iou3d_oracle.py, and the wrapper-level row masks are pinned
 * against the reference's own python (_check_coplanar/_check_nonzero executed
 * from /root/reference, fixtures in tests/golden/, generator
 * tests/golden/make_iou_golden.py).
 *
 * Build:  gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC  (no FMA
 * contraction: PyTorch3D's CPU build for baseline x86-64 has none either, and
 * the CUDA kernel is compiled with -fmad=false so face counts agree bit-exact).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define K_EPS 1e-8f   /* kEpsilon */
#define D_EPS 1e-3f   /* dEpsilon: coplanarity */
#define A_EPS 1e-4f   /* aEpsilon: area */
#define MAX_TRIS 512  /* oracle capacity per clipped list (asserted) */

typedef struct { float x, y, z; } v3;
typedef struct { v3 v[3]; } tri_t;
typedef struct { v3 v[4]; } face4_t;

static const int BOX_PLANES[6][4] = {
    {0, 1, 2, 3}, {3, 2, 6, 7}, {0, 1, 5, 4}, {0, 3, 7, 4}, {1, 2, 6, 5}, {4, 5, 6, 7}};
static const int BOX_TRIS[12][3] = {
    {0, 1, 2}, {0, 3, 2}, {4, 5, 6}, {4, 6, 7}, {1, 5, 6}, {1, 6, 2},
    {0, 4, 7}, {0, 7, 3}, {3, 2, 6}, {3, 6, 7}, {0, 1, 5}, {0, 4, 5}};

static inline v3 mk(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 add(v3 a, v3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 scl(float s, v3 a) { return mk(s * a.x, s * a.y, s * a.z); }
static inline v3 divs(v3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
static inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 cross(v3 a, v3 b) {
  return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float norm(v3 a) { return sqrtf(dot(a, a)); }

static inline v3 get_normal(v3 e0, v3 e1) {
  v3 n = cross(e0, e1);
  return divs(n, fmaxf(norm(n), K_EPS));
}
static inline v3 tri_center(const tri_t* t) {
  return divs(add(add(t->v[0], t->v[1]), t->v[2]), 3.0f);
}
static inline v3 plane_center(const face4_t* q) {
  return divs(add(add(add(q->v[0], q->v[1]), q->v[2]), q->v[3]), 4.0f);
}
/* unit normal of a triangle from the best-conditioned pair of (vertex - centre) */
static v3 tri_normal(const tri_t* t) {
  v3 ctr = tri_center(t);
  v3 n = mk(0.f, 0.f, 0.f);
  float max_dist = -1.0f;
  for (int i = 0; i < 2; ++i)
    for (int j = i + 1; j < 3; ++j) {
      v3 a = sub(t->v[i], ctr), b = sub(t->v[j], ctr);
      float dist = norm(cross(a, b));
      if (dist > max_dist) { max_dist = dist; n = get_normal(a, b); }
    }
  return n;
}
/* unit normal of a box face, flipped to point toward the box centre */
static v3 plane_normal_direction(const face4_t* q, v3 center) {
  v3 pc = plane_center(q);
  v3 n = mk(0.f, 0.f, 0.f);
  float max_dist = -1.0f;
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 4; ++j) {
      v3 a = sub(q->v[i], pc), b = sub(q->v[j], pc);
      float dist = norm(cross(a, b));
      if (dist > max_dist) { max_dist = dist; n = get_normal(a, b); }
    }
  float c = dot(sub(center, pc), n);
  if (c < 0.0f) n = scl(-1.0f, n);
  return n;
}
static inline float face_area(const tri_t* t) {
  v3 n = cross(sub(t->v[1], t->v[0]), sub(t->v[2], t->v[0]));
  return norm(n) / 2.0f;
}
static inline int is_inside(v3 pc, v3 n, v3 p) { return dot(sub(p, pc), n) >= 0.0f; }

static v3 plane_edge_intersection(v3 pc, v3 n, v3 p0, v3 p1) {
  v3 direc = sub(p1, p0);
  direc = divs(direc, fmaxf(norm(direc), K_EPS));
  v3 p = divs(add(p1, p0), 2.0f);
  if (fabsf(dot(direc, n)) >= D_EPS) {
    float top = -1.0f * dot(sub(p0, pc), n);
    float bot = dot(sub(p1, p0), n);
    float a = top / bot;
    p = add(p0, scl(a, sub(p1, p0)));
  }
  return p;
}
/* most distant vertex pair between two small vertex sets (first max wins) */
static void argmax_verts(const v3* a, int na, const v3* b, int nb, int* ia, int* ib) {
  float best = -1.0f; *ia = 0; *ib = 0;
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < nb; ++j) {
      v3 d = sub(a[i], b[j]);
      float dist = dot(d, d);
      if (dist > best) { best = dist; *ia = i; *ib = j; }
    }
}
static int is_coplanar_tri_plane(const tri_t* t, const face4_t* q, v3 n) {
  v3 nt = tri_normal(t);
  int check1 = fabsf(dot(nt, n)) > 1.0f - D_EPS;
  int i, j;
  argmax_verts(t->v, 3, q->v, 4, &i, &j);
  v3 d = sub(t->v[i], q->v[j]);
  d = divs(d, fmaxf(norm(d), K_EPS));
  int check2 = fabsf(dot(d, n)) < D_EPS;
  return check1 && check2;
}
static int is_coplanar_tri_tri(const tri_t* t1, const tri_t* t2) {
  v3 n1 = tri_normal(t1), n2 = tri_normal(t2);
  int check1 = fabsf(dot(n1, n2)) > 1.0f - D_EPS;
  int i, j;
  argmax_verts(t1->v, 3, t2->v, 3, &i, &j);
  v3 d = sub(t1->v[i], t2->v[j]);
  d = divs(d, fmaxf(norm(d), K_EPS));
  int check2 = (fabsf(dot(d, n1)) < D_EPS) || (fabsf(dot(d, n2)) < D_EPS);
  return check1 && check2;
}
static inline tri_t mktri(v3 a, v3 b, v3 c) { tri_t t; t.v[0] = a; t.v[1] = b; t.v[2] = c; return t; }

/* clip one triangle by one (inward-normal) plane; returns number of output tris (0..2) */
static int clip_tri_by_plane(const face4_t* q, v3 n, const tri_t* t, tri_t* out) {
  v3 v0 = t->v[0], v1 = t->v[1], v2 = t->v[2];
  if (is_coplanar_tri_plane(t, q, n)) { out[0] = *t; return 1; }
  v3 pc = plane_center(q);
  int in0 = is_inside(pc, n, v0), in1 = is_inside(pc, n, v1), in2 = is_inside(pc, n, v2);
  if (in0 && in1 && in2) { out[0] = *t; return 1; }
  if (!in0 && !in1 && !in2) return 0;
  v3 vout, vin1, vin2, vin, vout1, vout2;
  int one_out = 0;
  if (in0 && in1 && !in2) { one_out = 1; vout = v2; vin1 = v0; vin2 = v1; }
  else if (in0 && !in1 && in2) { one_out = 1; vout = v1; vin1 = v0; vin2 = v2; }
  else if (!in0 && in1 && in2) { one_out = 1; vout = v0; vin1 = v1; vin2 = v2; }
  else if (in0 && !in1 && !in2) { vin = v0; vout1 = v1; vout2 = v2; }
  else if (!in0 && !in1 && in2) { vin = v2; vout1 = v0; vout2 = v1; }
  else { vin = v1; vout1 = v0; vout2 = v2; }
  if (one_out) {
    v3 p1 = plane_edge_intersection(pc, n, vin1, vout);
    v3 p2 = plane_edge_intersection(pc, n, vin2, vout);
    out[0] = mktri(vin1, p1, vin2);
    out[1] = mktri(p1, p2, vin2);
    return 2;
  }
  v3 p1 = plane_edge_intersection(pc, n, vin, vout1);
  v3 p2 = plane_edge_intersection(pc, n, vin, vout2);
  out[0] = mktri(vin, p1, p2);
  return 1;
}

static int box_intersections(const tri_t* tris, int ntris, const face4_t* planes, v3 center,
                             tri_t* out /* MAX_TRIS */, int* overflow) {
  static __thread tri_t bufA[MAX_TRIS], bufB[MAX_TRIS];
  tri_t* cur = bufA; tri_t* nxt = bufB;
  memcpy(cur, tris, sizeof(tri_t) * ntris);
  int n = ntris;
  for (int p = 0; p < 6; ++p) {
    v3 nrm = plane_normal_direction(&planes[p], center);
    int m = 0;
    for (int t = 0; t < n; ++t) {
      tri_t o[2];
      int k = clip_tri_by_plane(&planes[p], nrm, &cur[t], o);
      for (int u = 0; u < k; ++u) {
        if (m >= MAX_TRIS) { *overflow = 1; break; }
        nxt[m++] = o[u];
      }
    }
    tri_t* s = cur; cur = nxt; nxt = s; n = m;
  }
  memcpy(out, cur, sizeof(tri_t) * n);
  return n;
}

static float box_volume(const tri_t* tris, int n, v3 c) {
  float vol = 0.0f;
  for (int t = 0; t < n; ++t) {
    v3 a = sub(tris[t].v[0], c), b = sub(tris[t].v[1], c), d = sub(tris[t].v[2], c);
    float area = dot(a, cross(b, d));
    vol = vol + fabsf(area) / 6.0f;
  }
  return vol;
}
static v3 polyhedron_center(const tri_t* tris, int n) {
  float x = 0, y = 0, z = 0;
  for (int t = 0; t < n; ++t) {
    x += (tris[t].v[0].x + tris[t].v[1].x + tris[t].v[2].x) / 3.0f;
    y += (tris[t].v[0].y + tris[t].v[1].y + tris[t].v[2].y) / 3.0f;
    z += (tris[t].v[0].z + tris[t].v[1].z + tris[t].v[2].z) / 3.0f;
  }
  return mk(x / n, y / n, z / n);
}
static void box_parts(const float* b, tri_t* tris, face4_t* planes, v3* center) {
  v3 c[8];
  for (int i = 0; i < 8; ++i) c[i] = mk(b[3 * i], b[3 * i + 1], b[3 * i + 2]);
  for (int t = 0; t < 12; ++t) for (int k = 0; k < 3; ++k) tris[t].v[k] = c[BOX_TRIS[t][k]];
  for (int p = 0; p < 6; ++p) for (int k = 0; k < 4; ++k) planes[p].v[k] = c[BOX_PLANES[p][k]];
  v3 s = mk(0.f, 0.f, 0.f);
  for (int i = 0; i < 8; ++i) s = add(s, c[i]);
  *center = divs(s, 8.0f);
}

/* one (box1, box2) pair: vol, iou, nfaces (triangles of the final polyhedron),
 * and the per-side clipped counts (n1 = box1-side, n2 = box2-side before de-dup). */
static void pair_iou(const float* b1, const float* b2, float* vol_o, float* iou_o,
                     int32_t* nfaces_o, int32_t* nside_o) {
  tri_t t1[12], t2[12]; face4_t p1[6], p2[6]; v3 c1, c2;
  box_parts(b1, t1, p1, &c1);
  box_parts(b2, t2, p2, &c2);
  static __thread tri_t i1[2 * MAX_TRIS], i2[MAX_TRIS];
  int ovf = 0;
  int n1 = box_intersections(t1, 12, p2, c2, i1, &ovf);
  int n2 = box_intersections(t2, 12, p1, c1, i2, &ovf);
  if (nside_o) { nside_o[0] = n1; nside_o[1] = n2; }
  int nf = n1;
  if (n2 > 0) {
    static __thread uint8_t keep[MAX_TRIS];
    memset(keep, 1, n2);
    for (int a = 0; a < n1; ++a)
      for (int b = 0; b < n2; ++b) {
        int cop = is_coplanar_tri_tri(&i1[a], &i2[b]);
        float area = face_area(&i1[a]);
        if (cop && area > A_EPS) keep[b] = 0;
      }
    for (int b = 0; b < n2; ++b) if (keep[b]) i1[nf++] = i2[b];
  }
  float vol = 0.0f, iou = 0.0f;
  if (nf > 0) {
    v3 pc = polyhedron_center(i1, nf);
    vol = box_volume(i1, nf, pc);
    float v1 = box_volume(t1, 12, c1), v2 = box_volume(t2, 12, c2);
    iou = vol / (v1 + v2 - vol);
  }
  *vol_o = vol; *iou_o = iou;
  if (nfaces_o) *nfaces_o = ovf ? -1 : nf;
}

/* row-parallel helper (pthreads; libgomp is absent in this image) */
typedef struct {
  const float *b1, *b2; int N, M, paired; float *vol, *iou; int32_t *nfaces, *nside;
  int tid, nthreads;
} job_t;
static void* job_run(void* arg) {
  job_t* J = (job_t*)arg;
  if (J->paired) {
    for (int k = J->tid; k < J->N; k += J->nthreads)
      pair_iou(J->b1 + 24 * (size_t)k, J->b2 + 24 * (size_t)k, J->vol + k, J->iou + k,
               J->nfaces ? J->nfaces + k : 0, 0);
    return 0;
  }
  for (int i = J->tid; i < J->N; i += J->nthreads)
    for (int j = 0; j < J->M; ++j) {
      size_t k = (size_t)i * J->M + j;
      pair_iou(J->b1 + 24 * (size_t)i, J->b2 + 24 * (size_t)j, J->vol + k, J->iou + k,
               J->nfaces ? J->nfaces + k : 0, J->nside ? J->nside + 2 * k : 0);
    }
  return 0;
}
static void run_jobs(job_t proto, int threads) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  if (threads == 1) { proto.tid = 0; proto.nthreads = 1; job_run(&proto); return; }
  pthread_t th[256]; job_t jobs[256];
  for (int t = 0; t < threads; ++t) {
    jobs[t] = proto; jobs[t].tid = t; jobs[t].nthreads = threads;
    pthread_create(&th[t], 0, job_run, &jobs[t]);
  }
  for (int t = 0; t < threads; ++t) pthread_join(th[t], 0);
}

/* == pytorch3d._C.iou_box3d(boxes1 (N,8,3), boxes2 (M,8,3)) -> vol (N,M), iou (N,M).
 * nfaces/nside may be NULL.  threads<=1: the reference's serial double loop. */
void oracle_iou_box3d(const float* boxes1, int N, const float* boxes2, int M, float* vol,
                      float* iou, int32_t* nfaces, int32_t* nside, int threads) {
  job_t J = {boxes1, boxes2, N, M, 0, vol, iou, nfaces, nside, 0, 1};
  run_jobs(J, threads);
}
/* paired mode: pair k = (boxes1[k], boxes2[k]) */
void oracle_iou_box3d_paired(const float* boxes1, const float* boxes2, int P, float* vol,
                             float* iou, int32_t* nfaces, int threads) {
  job_t J = {boxes1, boxes2, P, 1, 1, vol, iou, nfaces, 0, 0, 1};
  run_jobs(J, threads);
}

/* omni3d_evaluation.py:65-86 — note the reference SUMS the six per-face offsets
 * (mat1.bmm(mat2)) before abs(): |sum_p (v3-v0).n_p| < eps.  F.normalize eps 1e-12. */
static inline v3 f_normalize(v3 a) { return divs(a, fmaxf(norm(a), 1e-12f)); }
void oracle_check_boxes(const float* boxes, int N, float eps_coplanar, float eps_nonzero,
                        uint8_t* coplanar_ok, uint8_t* nonzero_ok) {
  for (int b = 0; b < N; ++b) {
    const float* B = boxes + 24 * (size_t)b;
    v3 c[8];
    for (int i = 0; i < 8; ++i) c[i] = mk(B[3 * i], B[3 * i + 1], B[3 * i + 2]);
    float acc = 0.0f;
    for (int p = 0; p < 6; ++p) {
      v3 v0 = c[BOX_PLANES[p][0]], v1 = c[BOX_PLANES[p][1]], v2 = c[BOX_PLANES[p][2]],
         v3_ = c[BOX_PLANES[p][3]];
      v3 e0 = f_normalize(sub(v1, v0)), e1 = f_normalize(sub(v2, v0));
      v3 n = f_normalize(cross(e0, e1));
      v3 d = sub(v3_, v0);
      acc += d.x * n.x; acc += d.y * n.y; acc += d.z * n.z;
    }
    coplanar_ok[b] = fabsf(acc) < eps_coplanar;
    int ok = 1;
    for (int t = 0; t < 12; ++t) {
      v3 v0 = c[BOX_TRIS[t][0]], v1 = c[BOX_TRIS[t][1]], v2 = c[BOX_TRIS[t][2]];
      v3 n = cross(sub(v1, v0), sub(v2, v0));
      float area = norm(n) / 2.0f;
      if (!(area > eps_nonzero)) ok = 0;
    }
    nonzero_ok[b] = (uint8_t)ok;
  }
}

/* omni3d_evaluation.py:106-166 box3d_overlap: iou with offending dt rows zeroed.
 * Returns counts of invalid rows through n_bad[0] (non-coplanar), n_bad[1] (zero-area). */
void oracle_box3d_overlap(const float* boxes_dt, int N, const float* boxes_gt, int M,
                          float eps_coplanar, float eps_nonzero, float* iou, float* vol_scratch,
                          int32_t* n_bad, int threads) {
  uint8_t* cop = (uint8_t*)malloc(N > 0 ? N : 1);
  uint8_t* nz = (uint8_t*)malloc(N > 0 ? N : 1);
  oracle_check_boxes(boxes_dt, N, eps_coplanar, eps_nonzero, cop, nz);
  oracle_iou_box3d(boxes_dt, N, boxes_gt, M, vol_scratch, iou, 0, 0, threads);
  int nb0 = 0, nb1 = 0;
  for (int i = 0; i < N; ++i) {
    if (!cop[i]) { ++nb0; for (int j = 0; j < M; ++j) iou[(size_t)i * M + j] = 0.0f; }
    if (!nz[i]) { ++nb1; for (int j = 0; j < M; ++j) iou[(size_t)i * M + j] = 0.0f; }
  }
  if (n_bad) { n_bad[0] = nb0; n_bad[1] = nb1; }
  free(cop); free(nz);
}
