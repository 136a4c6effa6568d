/* oracle/conv3p_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * A plain-C restatement of the reference CPU algorithm for the Conv3p hot path
 * (hkust-vgd/pointwise, tf_ops/conv3p/tf_conv3p_atrous.cpp; citations are file lines).  It is the
 * checker for the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may load it.  Parity status: PINNED -- tests/test_oracle.py checks every function here against
 * (a) the reference's own object code (oracle/_ref, built from /root/reference by oracle/Makefile)
 * when it is present and (b) the golden vectors under tests/golden/ that were generated from it.
 *
 * Build: gcc -O2 -ffp-contract=off (no -ffast-math): every float operation below must round
 * exactly like the reference's g++ -O3 build on x86-64 (SSE scalar fp32, no FMA contraction).
 *
 * Filter is fixed to 3x3x3 taps (all reference models, pointcnn2_acsd.py:50-66); strides are per
 * axis (x, y, z) as read at tf_conv3p_atrous.cpp:438-440.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TAPS 3
#define NCELL 27

/* ---------------------------------------------------------------------------------------------
 * Uniform grid over one cloud: bounding box, cell size = voxel size, counting sort of the point
 * indices by cell (stable in point index).  Follows Grid::Grid, tf_conv3p_atrous.cpp:157-227.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* xyz; /* [n,3] */
  int n;
  float cell_size;
  float lo[3], hi[3];
  int dim[3];
  int* start;  /* [cells+1] */
  int* member; /* [n] point indices ordered by cell */
} cloud_grid;

static int grid_coord(const cloud_grid* g, float v, int axis) {
  /* (int)((x - vmin) / radius), fp32 throughout -- :192-194, :251-253 */
  return (int)((v - g->lo[axis]) / g->cell_size);
}

static int grid_build(cloud_grid* g, const float* xyz, int n, float cell_size) {
  g->xyz = xyz;
  g->n = n;
  g->cell_size = cell_size;
  for (int a = 0; a < 3; ++a) { /* :163-164 */
    g->lo[a] = 1e6f;
    g->hi[a] = -1e6f;
  }
  for (int i = 0; i < n; ++i) /* :165-177 */
    for (int a = 0; a < 3; ++a) {
      float v = xyz[3 * i + a];
      if (v < g->lo[a]) g->lo[a] = v; /* std::min(vmin, x) */
      if (g->hi[a] < v) g->hi[a] = v; /* std::max(vmax, x) */
    }
  for (int a = 0; a < 3; ++a) /* :179-181, +2 padding */
    g->dim[a] = (int)((g->hi[a] - g->lo[a]) / cell_size) + 2;
  long long cells = (long long)g->dim[0] * g->dim[1] * g->dim[2];
  if (cells <= 0 || cells > (1LL << 28)) return -1;
  g->start = (int*)calloc((size_t)cells + 1, sizeof(int));
  g->member = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int* cell_of = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  if (!g->start || !g->member || !cell_of) return -1;
  for (int i = 0; i < n; ++i) { /* :187-203 */
    int cx = grid_coord(g, xyz[3 * i + 0], 0);
    int cy = grid_coord(g, xyz[3 * i + 1], 1);
    int cz = grid_coord(g, xyz[3 * i + 2], 2);
    cell_of[i] = (cz * g->dim[1] + cy) * g->dim[0] + cx;
    g->start[cell_of[i] + 1]++;
  }
  for (long long c = 0; c < cells; ++c) g->start[c + 1] += g->start[c]; /* :205-210 */
  int* fill = (int*)calloc((size_t)cells, sizeof(int));
  if (!fill) return -1;
  for (int i = 0; i < n; ++i) { /* :218-224, ascending i inside a cell */
    int c = cell_of[i];
    g->member[g->start[c] + fill[c]++] = i;
  }
  free(fill);
  free(cell_of);
  return 0;
}

static void grid_free(cloud_grid* g) {
  free(g->start);
  free(g->member);
}

/* The box of `full = (3-1)*stride+1` voxels centred on a query coordinate: bounds are formed in
 * double and rounded to float on assignment -- :240-245 (int * double * float, then T xmin = ...). */
static float box_lo(float centre, int full, float voxel) {
  return (float)((double)centre - full * 0.5 * (double)voxel);
}
static float box_hi(float centre, int full, float voxel) {
  return (float)((double)centre + full * 0.5 * (double)voxel);
}

/* Tap index of coordinate v inside a box starting at lo, or -1 for a dilation hole.
 * fp32 subtract, fp32 IEEE divide, truncation, clamp to full-1, hole test, /stride -- :280-288. */
static int tap_of(float v, float lo, float voxel, int full, int stride) {
  int c = (int)((v - lo) / voxel);
  if (c > full - 1) c = full - 1;
  if (c % stride != 0) return -1;
  return c / stride;
}

/* Enumerate the neighbours of query (qx,qy,qz) in the reference's emission order
 * (cells oz,oy,ox ascending, then cell membership order) -- Grid::neighbor, :232-301.
 * Writes point ids to out_j and kernel cells to out_f (either may be NULL), per-cell counts to
 * cell_count[27] (zeroed here, :257).  Returns the number of neighbours. */
static int grid_query(const cloud_grid* g, float qx, float qy, float qz, const int stride[3],
                      float voxel, int* out_j, int* out_f, int cell_count[NCELL]) {
  const float q[3] = {qx, qy, qz};
  int full[3], reach[3], centre[3];
  float lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    full[a] = (TAPS - 1) * stride[a] + 1;        /* :235-237 */
    lo[a] = box_lo(q[a], full[a], voxel);        /* :240-245 */
    hi[a] = box_hi(q[a], full[a], voxel);
    reach[a] = (int)((full[a] + 1) * 0.5);       /* :247-249 */
    centre[a] = grid_coord(g, q[a], a);          /* :251-253 */
  }
  memset(cell_count, 0, sizeof(int) * NCELL);
  int found = 0;
  for (int oz = -reach[2]; oz <= reach[2]; ++oz)   /* :260-266 */
    for (int oy = -reach[1]; oy <= reach[1]; ++oy)
      for (int ox = -reach[0]; ox <= reach[0]; ++ox) {
        int cx = centre[0] + ox, cy = centre[1] + oy, cz = centre[2] + oz;
        if (cx < 0 || cx >= g->dim[0] || cy < 0 || cy >= g->dim[1] || cz < 0 || cz >= g->dim[2])
          continue;
        int cell = (cz * g->dim[1] + cy) * g->dim[0] + cx;
        for (int m = g->start[cell]; m < g->start[cell + 1]; ++m) { /* :269-296 */
          int j = g->member[m];
          float vx = g->xyz[3 * j], vy = g->xyz[3 * j + 1], vz = g->xyz[3 * j + 2];
          if (vx < lo[0] || vx > hi[0] || vy < lo[1] || vy > hi[1] || vz < lo[2] || vz > hi[2])
            continue; /* closed box, :277 */
          int tx = tap_of(vx, lo[0], voxel, full[0], stride[0]);
          int ty = tap_of(vy, lo[1], voxel, full[1], stride[1]);
          int tz = tap_of(vz, lo[2], voxel, full[2], stride[2]);
          if (tx < 0 || ty < 0 || tz < 0) continue; /* hole, :285 */
          int f = (tz * TAPS + ty) * TAPS + tx;      /* :290 */
          if (out_j) out_j[found] = j;
          if (out_f) out_f[found] = f;
          cell_count[f]++;
          found++;
        }
      }
  return found;
}

/* ---------------------------------------------------------------------------------------------
 * Index-level entry points (one cloud).
 * ------------------------------------------------------------------------------------------- */

/* count[n,27]: neighbours of point i per kernel cell -- Grid::neighbor_count, :369-379. */
int oracle_neighbor_count_f32(const float* xyz, int n, const int stride[3], float voxel,
                              int* count) {
  cloud_grid g;
  if (grid_build(&g, xyz, n, voxel)) return -1;
  for (int i = 0; i < n; ++i)
    grid_query(&g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], stride, voxel, NULL, NULL,
               count + (size_t)i * NCELL);
  grid_free(&g);
  return 0;
}

/* (j, f) lists in reference emission order; off[n+1].  Returns total pairs (may exceed capacity;
 * the excess is counted, not stored). */
long long oracle_neighbors_f32(const float* xyz, int n, const int stride[3], float voxel,
                               long long* off, int* nbr_j, int* nbr_f, long long capacity) {
  cloud_grid g;
  if (grid_build(&g, xyz, n, voxel)) return -1;
  int* tj = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int* tf = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int cc[NCELL];
  long long total = 0;
  for (int i = 0; i < n; ++i) {
    off[i] = total;
    int k = grid_query(&g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], stride, voxel, tj, tf, cc);
    for (int m = 0; m < k; ++m, ++total)
      if (total < capacity) {
        nbr_j[total] = tj[m];
        nbr_f[total] = tf[m];
      }
  }
  off[n] = total;
  free(tj);
  free(tf);
  grid_free(&g);
  return total;
}

/* The same predicate with NO search structure: every point of the cloud is a candidate.  Used by
 * tests to show that the grid window never drops a neighbour (hard part 2). */
int oracle_neighbor_count_bruteforce_f32(const float* xyz, int n, const int stride[3], float voxel,
                                         int* count) {
  memset(count, 0, sizeof(int) * (size_t)n * NCELL);
  for (int i = 0; i < n; ++i) {
    int full[3];
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
      full[a] = (TAPS - 1) * stride[a] + 1;
      lo[a] = box_lo(xyz[3 * i + a], full[a], voxel);
      hi[a] = box_hi(xyz[3 * i + a], full[a], voxel);
    }
    for (int j = 0; j < n; ++j) {
      float vx = xyz[3 * j], vy = xyz[3 * j + 1], vz = xyz[3 * j + 2];
      if (vx < lo[0] || vx > hi[0] || vy < lo[1] || vy > hi[1] || vz < lo[2] || vz > hi[2]) continue;
      int tx = tap_of(vx, lo[0], voxel, full[0], stride[0]);
      int ty = tap_of(vy, lo[1], voxel, full[1], stride[1]);
      int tz = tap_of(vz, lo[2], voxel, full[2], stride[2]);
      if (tx < 0 || ty < 0 || tz < 0) continue;
      count[(size_t)i * NCELL + (tz * TAPS + ty) * TAPS + tx]++;
    }
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Forward -- Conv3pOp<CPU,T>::Compute, :401-509.  out[B,N,Cout] is zero-filled then accumulated
 * in the reference order: neighbours in emission order, c outer, k inner, one divide per term
 * (:480-494).  acc64 != NULL additionally returns the same sum accumulated in double, and
 * abs64 != NULL the sum of |terms| (tolerance scale for the CUDA path's different summation order).
 * ------------------------------------------------------------------------------------------- */
int oracle_conv3p_forward_f32(const float* points, const float* input, const float* filter,
                              const int stride[3], float voxel, int B, int N, int Cin, int Cout,
                              float* out, double* acc64, double* abs64) {
  int err = 0;
  memset(out, 0, sizeof(float) * (size_t)B * N * Cout); /* :451 */
  if (acc64) memset(acc64, 0, sizeof(double) * (size_t)B * N * Cout);
  if (abs64) memset(abs64, 0, sizeof(double) * (size_t)B * N * Cout);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (int b = 0; b < B; ++b) { /* :456 */
    const float* xyz = points + (size_t)b * N * 3;
    const float* in = input + (size_t)b * N * Cin;
    float* o = out + (size_t)b * N * Cout;
    cloud_grid g;
    if (grid_build(&g, xyz, N, voxel)) { /* :463 */
      err = -1;
      continue;
    }
    int* nj = (int*)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
    int* nf = (int*)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
    int cc[NCELL];
    for (int i = 0; i < N; ++i) { /* :473 */
      int k_i = grid_query(&g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], stride, voxel, nj, nf, cc);
      for (int m = 0; m < k_i; ++m) { /* :480 */
        const float* row = in + (size_t)nj[m] * Cin;
        int f = nf[m];
        float members = (float)cc[f]; /* (T)fsize, :492 */
        for (int c = 0; c < Cout; ++c)
          for (int k = 0; k < Cin; ++k) {
            float w = filter[((size_t)f * Cin + k) * Cout + c]; /* :490 */
            float term = w * row[k] / members;                  /* :492 */
            o[(size_t)i * Cout + c] += term;
            if (acc64) {
              size_t at = ((size_t)b * N + i) * Cout + c;
              double t = (double)w * (double)row[k] / (double)cc[f];
              acc64[at] += t;
              if (abs64) abs64[at] += fabs(t);
            }
          }
      }
    }
    free(nj);
    free(nf);
    grid_free(&g);
  }
  return err;
}

/* ---------------------------------------------------------------------------------------------
 * Backward -- Conv3pGradOp<CPU,T>::Compute, :526-720.  For every j and every ii in N(j) the cell
 * of j is recomputed in ii's frame WITHOUT a box test (:658-669); holes are skipped (:672);
 * count = table[ii, f'] and pairs with count == 0 are skipped (:678-679, "non-symmetric
 * neighbor issue").  grad_input[j,k] and grad_filter[f',k,c] accumulate in that order (:682-698).
 * The reference's per-thread grad_filter copies (:611-621, :709-716) only change the summation
 * order of grad_filter; here clouds are reduced in batch order.  Optional double outputs as above.
 * ------------------------------------------------------------------------------------------- */
int oracle_conv3p_backward_f32(const float* grad_out, const float* points, const float* input,
                               const float* filter, const int stride[3], float voxel, int B, int N,
                               int Cin, int Cout, float* grad_input, float* grad_filter,
                               double* gi64, double* gi_abs64, double* gf64, double* gf_abs64) {
  int err = 0;
  const size_t nw = (size_t)NCELL * Cin * Cout;
  memset(grad_input, 0, sizeof(float) * (size_t)B * N * Cin); /* :580 */
  memset(grad_filter, 0, sizeof(float) * nw);                 /* :590 */
  if (gi64) memset(gi64, 0, sizeof(double) * (size_t)B * N * Cin);
  if (gi_abs64) memset(gi_abs64, 0, sizeof(double) * (size_t)B * N * Cin);
  if (gf64) memset(gf64, 0, sizeof(double) * nw);
  if (gf_abs64) memset(gf_abs64, 0, sizeof(double) * nw);
  float* gf_cloud = (float*)calloc(nw * (size_t)(B > 0 ? B : 1), sizeof(float));
  double* gf64_cloud = gf64 ? (double*)calloc(nw * (size_t)(B > 0 ? B : 1), sizeof(double)) : NULL;
  double* gfa_cloud = gf_abs64 ? (double*)calloc(nw * (size_t)(B > 0 ? B : 1), sizeof(double)) : NULL;
  if (!gf_cloud) return -1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (int b = 0; b < B; ++b) { /* :622 */
    const float* xyz = points + (size_t)b * N * 3;
    const float* in = input + (size_t)b * N * Cin;
    const float* go = grad_out + (size_t)b * N * Cout;
    float* gi = grad_input + (size_t)b * N * Cin;
    float* gfb = gf_cloud + nw * b;
    cloud_grid g;
    if (grid_build(&g, xyz, N, voxel)) { /* :629 */
      err = -1;
      continue;
    }
    int* table = (int*)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1) * NCELL);
    int* nj = (int*)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
    int cc[NCELL];
    for (int i = 0; i < N; ++i) /* :637-639 */
      grid_query(&g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], stride, voxel, NULL, NULL,
                 table + (size_t)i * NCELL);
    int full[3];
    for (int a = 0; a < 3; ++a) full[a] = (TAPS - 1) * stride[a] + 1; /* :562-564 */
    for (int j = 0; j < N; ++j) { /* :647 */
      float x = xyz[3 * j], y = xyz[3 * j + 1], z = xyz[3 * j + 2];
      int k_j = grid_query(&g, x, y, z, stride, voxel, nj, NULL, cc); /* :652 */
      for (int m = 0; m < k_j; ++m) {                                 /* :654 */
        int ii = nj[m];
        float lx = box_lo(xyz[3 * ii], full[0], voxel); /* :658-664 */
        float ly = box_lo(xyz[3 * ii + 1], full[1], voxel);
        float lz = box_lo(xyz[3 * ii + 2], full[2], voxel);
        int tx = tap_of(x, lx, voxel, full[0], stride[0]); /* :667-675, no inside test */
        int ty = tap_of(y, ly, voxel, full[1], stride[1]);
        int tz = tap_of(z, lz, voxel, full[2], stride[2]);
        if (tx < 0 || ty < 0 || tz < 0) continue;
        int f = (tz * TAPS + ty) * TAPS + tx; /* :677 */
        int members = table[(size_t)ii * NCELL + f];
        if (members == 0) continue; /* :679 */
        float fm = (float)members;
        for (int c = 0; c < Cout; ++c)
          for (int k = 0; k < Cin; ++k) { /* :682-697 */
            size_t wi = ((size_t)f * Cin + k) * Cout + c;
            float up = go[(size_t)ii * Cout + c];
            gi[(size_t)j * Cin + k] += up * filter[wi] / fm;          /* :692 */
            gfb[wi] += up * in[(size_t)j * Cin + k] / fm;             /* :696 */
            if (gi64) {
              size_t at = ((size_t)b * N + j) * Cin + k;
              double t = (double)up * (double)filter[wi] / (double)members;
              gi64[at] += t;
              if (gi_abs64) gi_abs64[at] += fabs(t);
            }
            if (gf64_cloud) {
              double t = (double)up * (double)in[(size_t)j * Cin + k] / (double)members;
              gf64_cloud[nw * b + wi] += t;
              if (gfa_cloud) gfa_cloud[nw * b + wi] += fabs(t);
            }
          }
      }
    }
    free(table);
    free(nj);
    grid_free(&g);
  }
  for (int b = 0; b < B; ++b) /* :709-716 restated per cloud */
    for (size_t w = 0; w < nw; ++w) {
      grad_filter[w] += gf_cloud[nw * b + w];
      if (gf64) gf64[w] += gf64_cloud[nw * b + w];
      if (gf_abs64) gf_abs64[w] += gfa_cloud[nw * b + w];
    }
  free(gf_cloud);
  free(gf64_cloud);
  free(gfa_cloud);
  return err;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
