/*
 * cudecomp_oracle.c -- CPU restatement of the reference's pencil transpose and halo exchange.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product (cudecomp_b200/, libcudecomp.so) may include, link or call
 * this file. It is used by tests/ as the checker, by __graft_entry__.smoke() as the checker, and by
 * bench.py's cpu_baseline / --impl reference leg as the thing timed on the host cores.
 *
 * Parity status: PINNED. The reference cannot be compiled here (no MPI, cuTENSOR, NVHPC),
 * so this file restates its algorithm and is pinned by tests/test_oracle_golden.py against every known-answer
 * fixture the reference's own tests hold for this path: the golden pencil-info tables
 * (tests/ctest/api_tests.cc:92-153), the golden shifted-rank tables (api_tests.cc:1386-1432) and the analytic
 * global-index pattern with which the reference checks every transpose and halo result exactly
 * (tests/ctest/transpose_tests.cc:323-378, tests/ctest/halo_tests.cc:229-272).
 *
 * The reference has no host-memory path (its MPI arms hand device pointers to CUDA-aware MPI,
 * include/internal/comm_routines.h:325-413). This restatement keeps the reference's three phases and its wire
 * format, with all R ranks living in one process:
 *   pack     per destination rank, strided sub-block -> contiguous send block   (include/internal/transpose.h:533-617)
 *   exchange send block i of rank r -> receive block r of rank i (what MPI_Alltoallv does; comm_routines.h:363-413)
 *   unpack   per source rank, contiguous receive block -> strided sub-block, permuting when the source and
 *            destination memory orders differ                                    (transpose.h:650-895)
 * The reference permutes either while packing or while unpacking depending on the layout (transpose.h:426,651);
 * the result is the same permutation of the same bytes, so one variant (permute on unpack) is restated.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int32_t gdims[3];
  int32_t gdims_dist[3]; /* must be resolved (no zeros) */
  int32_t pdims[2];
  int32_t col_major;
  int32_t order[3][3]; /* [axis][memory position] -> global axis */
} oracle_grid_t;

typedef struct {
  int32_t shape[3], lo[3], hi[3], order[3]; /* by memory position */
  int32_t halo[3], pad[3];                  /* by global axis */
  int64_t size;
} oracle_pencil_t;

static int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }

/* Staging regions persist between calls, like the reference's workspace does (the caller allocates `work` once,
 * include/cudecomp.h:433-448): a timed run must not pay for fresh page faults on every transpose. */
#define ORACLE_MAX_STAGE 1024
static char* g_stage[ORACLE_MAX_STAGE];
static size_t g_stage_bytes[ORACLE_MAX_STAGE];

static char* stage_buffer(int slot, size_t bytes) {
  if (slot < 0 || slot >= ORACLE_MAX_STAGE) return NULL;
  if (g_stage_bytes[slot] < bytes) {
    free(g_stage[slot]);
    g_stage[slot] = (char*)malloc(bytes ? bytes : 1);
    g_stage_bytes[slot] = g_stage[slot] ? bytes : 0;
    if (g_stage[slot]) memset(g_stage[slot], 0, bytes); /* touch the pages once */
  }
  return g_stage[slot];
}

void oracle_release(void) {
  for (int i = 0; i < ORACLE_MAX_STAGE; ++i) {
    free(g_stage[i]);
    g_stage[i] = NULL;
    g_stage_bytes[i] = 0;
  }
}

/* reference include/internal/common.h:318-331 */
static void pidx_of_rank(const oracle_grid_t* g, int rank, int pidx[2]) {
  if (g->col_major) {
    pidx[0] = rank % g->pdims[0];
    pidx[1] = rank / g->pdims[0];
  } else {
    pidx[0] = rank / g->pdims[1];
    pidx[1] = rank % g->pdims[1];
  }
}

/* reference include/internal/common.h:334-346 */
static int global_rank(const oracle_grid_t* g, const int pidx[2], int comm_is_row, int axis_rank) {
  if (g->col_major) return comm_is_row ? pidx[0] + axis_rank * g->pdims[0] : g->pdims[0] * pidx[1] + axis_rank;
  return comm_is_row ? g->pdims[1] * pidx[0] + axis_rank : pidx[1] + axis_rank * g->pdims[1];
}

/* reference src/cudecomp.cc:1317-1379 */
int oracle_pencil_info(const oracle_grid_t* g, int rank, int axis, const int32_t* halo, const int32_t* pad,
                       oracle_pencil_t* p) {
  int pidx[2], inv[3];
  pidx_of_rank(g, rank, pidx);
  for (int i = 0; i < 3; ++i) {
    p->order[i] = g->order[axis][i];
    inv[p->order[i]] = i;
  }
  int j = 0;
  p->size = 1;
  for (int i = 0; i < 3; ++i) {
    int ord = inv[i];
    if (i != axis) {
      int64_t d = g->gdims_dist[i] / g->pdims[j];
      int64_t mod = g->gdims_dist[i] % g->pdims[j];
      int64_t shape = d + (pidx[j] < mod ? 1 : 0);
      if (pidx[j] == min64(g->pdims[j], g->gdims_dist[i]) - 1) shape += g->gdims[i] - g->gdims_dist[i];
      p->shape[ord] = (int32_t)shape;
      p->lo[ord] = (int32_t)(pidx[j] * d + min64(pidx[j], mod));
      j++;
    } else {
      p->shape[ord] = g->gdims[i];
      p->lo[ord] = 0;
    }
    p->hi[ord] = p->lo[ord] + p->shape[ord] - 1;
    p->halo[i] = halo ? halo[i] : 0;
    p->pad[i] = pad ? pad[i] : 0;
    if (p->halo[i] < 0 || p->pad[i] < 0) return 1;
    p->shape[ord] += 2 * p->halo[i] + p->pad[i];
    p->size *= p->shape[ord];
  }
  return 0;
}

/* reference include/internal/common.h:579-589 */
static void get_splits(int64_t N, int n, int64_t pad, int64_t* splits, int64_t* offsets) {
  for (int i = 0; i < n; ++i) splits[i] = N / n + (i < N % n ? 1 : 0);
  splits[min64(N, n) - 1] += pad;
  offsets[0] = 0;
  for (int i = 1; i < n; ++i) offsets[i] = offsets[i - 1] + splits[i - 1];
}

/* reference include/internal/common.h:620-631 */
int oracle_has_empty_pencils(const oracle_grid_t* g, int axis) {
  int j = 0;
  for (int i = 0; i < 3; ++i) {
    if (i == axis) continue;
    if (g->gdims_dist[i] / g->pdims[j] == 0) return 1;
    j++;
  }
  return 0;
}

/* reference include/internal/common.h:349-366 */
static int64_t global_max_pencil_size(const oracle_grid_t* g, int axis) {
  int64_t size = 1;
  int j = 0;
  for (int i = 0; i < 3; ++i) {
    if (i != axis) {
      int64_t dim = (g->gdims_dist[i] + g->pdims[j] - 1) / g->pdims[j];
      dim += g->gdims[i] - g->gdims_dist[i];
      size *= dim;
      j++;
    } else {
      size *= g->gdims[i];
    }
  }
  return size;
}

/* reference include/internal/common.h:636-640 (256 bytes of 4-byte units) */
static int64_t align_count(int64_t count) { return (count * 4 + 255) / 256 * 256 / 4; }

/* reference src/cudecomp.cc:1411-1432 */
int64_t oracle_transpose_workspace_size(const oracle_grid_t* g) {
  int64_t x = global_max_pencil_size(g, 0), y = global_max_pencil_size(g, 1), z = global_max_pencil_size(g, 2);
  int64_t w[4] = {align_count(x) + y, align_count(y) + x, align_count(y) + z, align_count(z) + y};
  int64_t m = w[0];
  for (int i = 1; i < 4; ++i)
    if (w[i] > m) m = w[i];
  return m;
}

static void shape_g(const oracle_pencil_t* p, int64_t s[3]) {
  for (int i = 0; i < 3; ++i) s[p->order[i]] = p->shape[i];
}

/* reference src/cudecomp.cc:1434-1459 */
int64_t oracle_halo_workspace_size(const oracle_grid_t* g, int rank, int axis, const int32_t* halo) {
  oracle_pencil_t p;
  int64_t s[3], m = 0;
  oracle_pencil_info(g, rank, axis, halo, NULL, &p);
  shape_g(&p, s);
  for (int d = 0; d < 3; ++d) {
    int64_t v = 4 * align_count(s[(d + 1) % 3] * s[(d + 2) % 3] * p.halo[d]);
    if (v > m) m = v;
  }
  return m;
}

/* reference src/cudecomp.cc:1710-1755 */
int oracle_shifted_rank(const oracle_grid_t* g, int rank, int axis, int dim, int displacement, int periodic) {
  if (displacement == 0) return rank;
  if (dim == axis) return periodic ? rank : -1;
  int count = 0;
  for (int i = 0; i < 3; ++i) {
    if (i == axis) continue;
    if (i == dim) break;
    count++;
  }
  int comm_is_row = (count != 0);
  int pidx[2];
  pidx_of_rank(g, rank, pidx);
  int n = g->pdims[comm_is_row ? 1 : 0];
  int comm_rank = pidx[comm_is_row ? 1 : 0];
  int shifted = comm_rank + displacement;
  if (!periodic && (shifted < 0 || shifted >= n)) return -1;
  int peer = ((shifted % n) + n) % n;
  return global_rank(g, pidx, comm_is_row, peer);
}

/* element offset of local index lx (global-axis order) -- reference common.h:369-373 */
static int64_t ptr_offset(const oracle_pencil_t* p, const int64_t lx[3]) {
  return lx[p->order[0]] + lx[p->order[1]] * (int64_t)p->shape[0] +
         lx[p->order[2]] * (int64_t)p->shape[0] * (int64_t)p->shape[1];
}

/* element strides by global axis */
static void strides_g(const oracle_pencil_t* p, int64_t s[3]) {
  s[p->order[0]] = 1;
  s[p->order[1]] = p->shape[0];
  s[p->order[2]] = (int64_t)p->shape[0] * p->shape[1];
}

/*
 * Copy a 3-D block. Iteration runs in the order `ord` (ord[0] fastest); ext/sstr/dstr are by global axis.
 * Rows along ord[0] are moved with memcpy when both sides are contiguous there.
 */
static void copy_block(const char* src, char* dst, const int ord[3], const int64_t ext[3], const int64_t sstr[3],
                       const int64_t dstr[3], int es, int parallel) {
  const int a0 = ord[0], a1 = ord[1], a2 = ord[2];
  const int64_t n0 = ext[a0], n1 = ext[a1], n2 = ext[a2];
  const int contiguous = (sstr[a0] == 1 && dstr[a0] == 1);
  const int64_t rows = n1 * n2;
  if (n0 <= 0 || rows <= 0) return;
#pragma omp parallel for schedule(static) if (parallel && rows * n0 * es > (1 << 16))
  for (int64_t r = 0; r < rows; ++r) {
    const int64_t i1 = r % n1, i2 = r / n1;
    const char* s = src + (i1 * sstr[a1] + i2 * sstr[a2]) * es;
    char* d = dst + (i1 * dstr[a1] + i2 * dstr[a2]) * es;
    if (contiguous) {
      memcpy(d, s, (size_t)(n0 * es));
    } else {
      for (int64_t i0 = 0; i0 < n0; ++i0) memcpy(d + i0 * dstr[a0] * es, s + i0 * sstr[a0] * es, (size_t)es);
    }
  }
}

/* (ax, dir) -> a, b, c and communicator (reference include/internal/transpose.h:222-227) */
static void transpose_axes(int ax, int dir, int* a, int* b, int* c, int* comm_is_row) {
  *a = ax;
  *b = (dir > 0 ? ax + 1 : ax + 2) % 3;
  *c = (dir > 0 ? ax + 2 : ax + 1) % 3;
  *comm_is_row = (*a == 2 || *b == 2);
}

/*
 * One transpose over all nranks in-process ranks. in[r] / out[r] are rank r's pencils (may alias: in place).
 * Returns 0 ok, 2 unsupported (empty pencils), 1 bad arguments, 3 out of memory.
 */
int oracle_transpose(const oracle_grid_t* g, int ax, int dir, int es, void* const* in, void* const* out,
                     const int32_t* in_halo, const int32_t* out_halo, const int32_t* in_pad, const int32_t* out_pad) {
  const int nranks = g->pdims[0] * g->pdims[1];
  int a, b, c, comm_is_row;
  transpose_axes(ax, dir, &a, &b, &c, &comm_is_row);
  if (oracle_has_empty_pencils(g, a) || oracle_has_empty_pencils(g, b)) return 2;
  const int P = g->pdims[comm_is_row ? 1 : 0];

  int64_t* splits_a = (int64_t*)malloc(sizeof(int64_t) * 4 * (size_t)P);
  int64_t* off_a = splits_a + P;
  int64_t* splits_b = off_a + P;
  int64_t* off_b = splits_b + P;
  get_splits(g->gdims_dist[a], P, g->gdims[a] - g->gdims_dist[a], splits_a, off_a);
  get_splits(g->gdims_dist[b], P, g->gdims[b] - g->gdims_dist[b], splits_b, off_b);

  /* staging: send (packed A pencil) and receive (packed B pencil) regions of every rank */
  char** sendb = (char**)calloc((size_t)nranks * 2, sizeof(char*));
  char** recvb = sendb + nranks;
  int rc = 0;
  for (int r = 0; r < nranks && !rc; ++r) {
    oracle_pencil_t pa, pb;
    if (oracle_pencil_info(g, r, a, NULL, NULL, &pa) || oracle_pencil_info(g, r, b, NULL, NULL, &pb)) rc = 1;
    if (2 * nranks > ORACLE_MAX_STAGE) {
      rc = 3;
      break;
    }
    sendb[r] = stage_buffer(2 * r, (size_t)(pa.size > 0 ? pa.size : 1) * es);
    recvb[r] = stage_buffer(2 * r + 1, (size_t)(pb.size > 0 ? pb.size : 1) * es);
    if (!sendb[r] || !recvb[r]) rc = 3;
  }

  /* phase 1: pack */
  for (int r = 0; r < nranks && !rc; ++r) {
    oracle_pencil_t pa, pa_h;
    oracle_pencil_info(g, r, a, NULL, NULL, &pa);
    if (oracle_pencil_info(g, r, a, in_halo, in_pad, &pa_h)) {
      rc = 1;
      break;
    }
    int64_t sg[3], sstr[3];
    shape_g(&pa, sg);
    strides_g(&pa_h, sstr);
    for (int i = 0; i < P; ++i) {
      int64_t ext[3], lx[3], dstr[3];
      ext[a] = splits_a[i];
      ext[b] = sg[b];
      ext[c] = sg[c];
      lx[a] = off_a[i] + pa_h.halo[a];
      lx[b] = pa_h.halo[b];
      lx[c] = pa_h.halo[c];
      /* dense block in source memory order */
      int64_t acc = 1;
      for (int k = 0; k < 3; ++k) {
        dstr[pa.order[k]] = acc;
        acc *= ext[pa.order[k]];
      }
      const int64_t send_offset = off_a[i] * sg[b] * sg[c];
      copy_block((const char*)in[r] + ptr_offset(&pa_h, lx) * es, sendb[r] + send_offset * es, pa.order, ext, sstr,
                 dstr, es, 1);
    }
  }

  /* phase 2: all-to-all inside each row / column communicator */
  for (int r = 0; r < nranks && !rc; ++r) {
    int pidx[2];
    pidx_of_rank(g, r, pidx);
    const int me = pidx[comm_is_row ? 1 : 0];
    oracle_pencil_t pa;
    oracle_pencil_info(g, r, a, NULL, NULL, &pa);
    int64_t sga[3];
    shape_g(&pa, sga);
    for (int i = 0; i < P; ++i) {
      const int dst = global_rank(g, pidx, comm_is_row, i);
      oracle_pencil_t pbd;
      oracle_pencil_info(g, dst, b, NULL, NULL, &pbd);
      int64_t sgb[3];
      shape_g(&pbd, sgb);
      const int64_t send_offset = off_a[i] * sga[b] * sga[c];
      const int64_t send_count = splits_a[i] * sga[b] * sga[c];
      const int64_t recv_offset = off_b[me] * sgb[a] * sgb[c];
      /* split large messages over the host threads */
      const int64_t bytes = send_count * es;
      const int64_t chunk = 1 << 22;
      const int64_t nchunks = (bytes + chunk - 1) / chunk;
#pragma omp parallel for schedule(static) if (nchunks > 1)
      for (int64_t k = 0; k < nchunks; ++k) {
        const int64_t lo = k * chunk, len = (lo + chunk <= bytes) ? chunk : bytes - lo;
        memcpy(recvb[dst] + recv_offset * es + lo, sendb[r] + send_offset * es + lo, (size_t)len);
      }
    }
  }

  /* phase 3: unpack (permuting when the orders differ) */
  for (int r = 0; r < nranks && !rc; ++r) {
    int pidx[2];
    pidx_of_rank(g, r, pidx);
    const int me = pidx[comm_is_row ? 1 : 0];
    oracle_pencil_t pa, pb, pb_h;
    oracle_pencil_info(g, r, a, NULL, NULL, &pa);
    oracle_pencil_info(g, r, b, NULL, NULL, &pb);
    if (oracle_pencil_info(g, r, b, out_halo, out_pad, &pb_h)) {
      rc = 1;
      break;
    }
    int64_t sgb[3], dstr[3];
    shape_g(&pb, sgb);
    strides_g(&pb_h, dstr);
    for (int j = 0; j < P; ++j) {
      /* block from source j: a-range = mine (splits_a[me]), b-range = j's, dense in SOURCE memory order */
      int64_t ext[3], lx[3], sstr[3];
      ext[a] = splits_a[me];
      ext[b] = splits_b[j];
      ext[c] = sgb[c];
      int64_t acc = 1;
      for (int k = 0; k < 3; ++k) {
        sstr[pa.order[k]] = acc;
        acc *= ext[pa.order[k]];
      }
      lx[a] = pb_h.halo[a];
      lx[b] = off_b[j] + pb_h.halo[b];
      lx[c] = pb_h.halo[c];
      const int64_t recv_offset = off_b[j] * sgb[a] * sgb[c];
      /* iterate in destination order so that writes are contiguous */
      copy_block(recvb[r] + recv_offset * es, (char*)out[r] + ptr_offset(&pb_h, lx) * es, pb.order, ext, sstr, dstr,
                 es, 1);
    }
  }

  free(sendb);
  free(splits_a);
  return rc;
}

/*
 * Halo update of dimension `dim` of `ax`-pencils over all in-process ranks (reference include/internal/halo.h:40-315):
 * every rank packs its two boundary faces, the faces are exchanged with the +-1 neighbours
 * (include/internal/comm_routines.h:633-773), then unpacked into the halo cells. Faces span the halo-inclusive,
 * padding-exclusive extent of the other two dimensions. Returns 0 ok, 1 halo too wide / bad arguments, 2 empty pencils.
 */
int oracle_halo(const oracle_grid_t* g, int ax, int dim, int es, void* const* data, const int32_t* halo,
                const int32_t* periods, const int32_t* pad) {
  const int nranks = g->pdims[0] * g->pdims[1];
  if (oracle_has_empty_pencils(g, ax)) return 2;
  const int hw = halo ? halo[dim] : 0;
  if (hw == 0) return 0;
  const int periodic = periods ? periods[dim] : 0;

  char** faces = (char**)calloc((size_t)nranks * 2, sizeof(char*)); /* [r*2+side]: packed left/right interior face */
  int64_t* face_elems = (int64_t*)calloc((size_t)nranks, sizeof(int64_t));
  int rc = 0;

  /* validity: the face must come from the nearest neighbour alone (halo.h:120-145) */
  if (dim != ax) {
    int count = 0;
    for (int i = 0; i < 3; ++i) {
      if (i == ax) continue;
      if (i == dim) break;
      count++;
    }
    const int P = g->pdims[count == 0 ? 0 : 1];
    if (P > 1) {
      int64_t* sp = (int64_t*)malloc(sizeof(int64_t) * 2 * (size_t)P);
      get_splits(g->gdims_dist[dim], P, g->gdims[dim] - g->gdims_dist[dim], sp, sp + P);
      for (int i = 0; i < P; ++i)
        if (hw > sp[i]) rc = 1;
      free(sp);
    }
  }

  /* pack (halo.h:195-236) */
  for (int r = 0; r < nranks && !rc; ++r) {
    oracle_pencil_t ph, php;
    if (oracle_pencil_info(g, r, ax, halo, NULL, &ph) || oracle_pencil_info(g, r, ax, halo, pad, &php)) {
      rc = 1;
      break;
    }
    int64_t sh[3], shp[3], sstr[3], ext[3], dstr[3];
    shape_g(&ph, sh);
    shape_g(&php, shp);
    strides_g(&php, sstr);
    for (int k = 0; k < 3; ++k) ext[k] = (k == dim) ? hw : sh[k];
    int64_t acc = 1;
    for (int k = 0; k < 3; ++k) {
      dstr[php.order[k]] = acc;
      acc *= ext[php.order[k]];
    }
    face_elems[r] = acc;
    for (int side = 0; side < 2; ++side) {
      int64_t lx[3] = {0, 0, 0};
      lx[dim] = (side == 0) ? hw : shp[dim] - 2 * hw - php.pad[dim];
      faces[r * 2 + side] = (char*)malloc((size_t)(acc > 0 ? acc : 1) * es);
      copy_block((const char*)data[r] + ptr_offset(&php, lx) * es, faces[r * 2 + side], php.order, ext, sstr, dstr, es,
                 0);
    }
  }

  /* exchange + unpack (halo.h:238-276): my left halo <- left neighbour's right face, my right halo <- right
   * neighbour's left face */
  for (int r = 0; r < nranks && !rc; ++r) {
    oracle_pencil_t ph, php;
    oracle_pencil_info(g, r, ax, halo, NULL, &ph);
    oracle_pencil_info(g, r, ax, halo, pad, &php);
    int64_t sh[3], shp[3], dstr[3], ext[3], sstr[3];
    shape_g(&ph, sh);
    shape_g(&php, shp);
    strides_g(&php, dstr);
    for (int k = 0; k < 3; ++k) ext[k] = (k == dim) ? hw : sh[k];
    int64_t acc = 1;
    for (int k = 0; k < 3; ++k) {
      sstr[php.order[k]] = acc;
      acc *= ext[php.order[k]];
    }
    const int nb[2] = {oracle_shifted_rank(g, r, ax, dim, -1, periodic), oracle_shifted_rank(g, r, ax, dim, 1, periodic)};
    for (int side = 0; side < 2; ++side) {
      if (nb[side] < 0) continue; /* non-periodic boundary: halo cells stay as they are */
      int64_t lx[3] = {0, 0, 0};
      lx[dim] = (side == 0) ? 0 : shp[dim] - hw - php.pad[dim];
      const char* src = faces[nb[side] * 2 + (side == 0 ? 1 : 0)];
      copy_block(src, (char*)data[r] + ptr_offset(&php, lx) * es, php.order, ext, sstr, dstr, es, 0);
    }
  }

  for (int i = 0; i < nranks * 2; ++i) free(faces[i]);
  free(faces);
  free(face_elems);
  return rc;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
