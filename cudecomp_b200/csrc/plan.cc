// See plan.h.
#include "plan.h"

#include <algorithm>

#include "errors.h"

namespace cdb {

static bool same3(const int32_t* a, const int32_t* b) {
  for (int i = 0; i < 3; ++i)
    if ((a ? a[i] : 0) != (b ? b[i] : 0)) return false;
  return true;
}

static int64_t dot3(const std::array<int64_t, 3>& a, const std::array<int64_t, 3>& b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

TransposePlan buildTransposePlan(const GridGeom& g, const std::array<int, 2>& pidx, int ax, int dir,
                                 const int32_t in_halo[3], const int32_t out_halo[3], const int32_t in_pad[3],
                                 const int32_t out_pad[3], DstKind kind, bool inplace) {
  TransposePlan plan;
  plan.axes = transposeAxes(ax, dir);
  const int a = plan.axes.a, b = plan.axes.b, c = plan.axes.c;
  const int ci = plan.axes.comm; // pdims/pidx slot of the communicator

  // geometry first, so bad halo/padding arguments are reported before anything else (reference order,
  // transpose.h:248-259)
  const Pencil pa = pencilInfo(g, pidx, a, nullptr, nullptr);
  const Pencil pa_h = pencilInfo(g, pidx, a, in_halo, in_pad);
  const Pencil pb = pencilInfo(g, pidx, b, nullptr, nullptr);
  const Pencil pb_h = pencilInfo(g, pidx, b, out_halo, out_pad);
  if (hasEmptyPencils(g, a) || hasEmptyPencils(g, b))
    THROW_NOT_SUPPORTED("transposes on configurations with empty pencils not supported");

  const int P = g.pdims[ci];
  const int me = pidx[ci];
  plan.comm_size = P;
  plan.me = me;
  for (int i = 0; i < P; ++i) {
    auto pp = pidx;
    pp[ci] = i;
    plan.group_world.push_back(rankOfPidx(g, pp));
  }

  const auto splits_a = getSplits(g.gdims_dist[a], P, g.gdims[a] - g.gdims_dist[a]);
  const auto splits_b = getSplits(g.gdims_dist[b], P, g.gdims[b] - g.gdims_dist[b]);
  const auto off_a = prefixOffsets(splits_a);
  const auto off_b = prefixOffsets(splits_b);

  const bool orders_equal = (pa.order == pb.order);
  const bool halos_equal = same3(in_halo, out_halo) && same3(in_pad, out_pad);
  if (P == 1 && inplace && orders_equal && halos_equal) {
    plan.noop = true; // reference transpose.h:326-341
    return plan;
  }

  const auto shape_a = pa.shapeG();
  if (shape_a[b] != splits_b[me]) THROW_INTERNAL_ERROR("pencil/split mismatch");
  const auto sstr = pa_h.strideG();
  plan.src_elems = pa.size;

  for (int i = 0; i < P; ++i) {
    auto pp = pidx;
    pp[ci] = i;
    BoxDesc bx;
    bx.peer = i;
    bx.peer_world = plan.group_world[i];
    bx.ext[a] = splits_a[i];
    bx.ext[b] = shape_a[b];
    bx.ext[c] = shape_a[c];
    std::array<int64_t, 3> s0{}, d0{};
    s0[a] = off_a[i] + pa_h.halo[a];
    s0[b] = pa_h.halo[b];
    s0[c] = pa_h.halo[c];
    bx.sstr = sstr;
    bx.src_off = dot3(s0, sstr);
    if (kind == DstKind::FINAL) {
      const Pencil q = pencilInfo(g, pp, b, out_halo, out_pad);
      bx.dstr = q.strideG();
      d0[a] = q.halo[a];
      d0[b] = off_b[me] + q.halo[b];
      d0[c] = q.halo[c];
    } else {
      const Pencil q = pencilInfo(g, pp, b, nullptr, nullptr);
      bx.dstr = q.strideG();
      d0[b] = off_b[me];
    }
    bx.dst_off = dot3(d0, bx.dstr);
    if (i != me) plan.wire_elems += bx.count();
    plan.push.push_back(bx);
  }

  if (kind == DstKind::STAGE) {
    BoxDesc u;
    u.peer = me;
    u.peer_world = plan.group_world[me];
    auto sb = pb.shapeG();
    for (int k = 0; k < 3; ++k) u.ext[k] = sb[k];
    u.sstr = pb.strideG();
    u.dstr = pb_h.strideG();
    std::array<int64_t, 3> d0{pb_h.halo[0], pb_h.halo[1], pb_h.halo[2]};
    u.dst_off = dot3(d0, u.dstr);
    plan.unpack.push_back(u);
  }
  return plan;
}

TransposePlan buildPullTransposePlan(const GridGeom& g, const std::array<int, 2>& pidx, int ax, int dir,
                                     const int32_t in_halo[3], const int32_t out_halo[3], const int32_t in_pad[3],
                                     const int32_t out_pad[3], DstKind kind, bool inplace) {
  TransposePlan plan = buildTransposePlan(g, pidx, ax, dir, in_halo, out_halo, in_pad, out_pad, kind, inplace);
  if (plan.noop) return plan;
  const int ci = plan.axes.comm;
  plan.push.clear();
  plan.wire_elems = 0;
  for (int j = 0; j < plan.comm_size; ++j) {
    auto pj = pidx;
    pj[ci] = j;
    const TransposePlan theirs = buildTransposePlan(g, pj, ax, dir, in_halo, out_halo, in_pad, out_pad, kind, false);
    BoxDesc b = theirs.push[plan.me]; // what rank j sends to me: its source strides, my destination strides
    b.peer = j;
    b.peer_world = plan.group_world[j];
    if (j != plan.me) plan.wire_elems += b.count();
    plan.push.push_back(b);
  }
  return plan;
}

namespace {
// chunk k of [0, n) cut into K parts
inline std::pair<int64_t, int64_t> chunkRange(int64_t n, int K, int k) { return {k * n / K, (k + 1) * n / K}; }
} // namespace

PipelinedPlan buildPipelinedTransposePlan(const GridGeom& g, const std::array<int, 2>& pidx, int ax, int dir,
                                          const int32_t in_halo[3], const int32_t out_halo[3], const int32_t in_pad[3],
                                          const int32_t out_pad[3], bool inplace, int nchunks, bool pull, int elem_bytes,
                                          int64_t min_row_bytes) {
  if (pull) {
    PipelinedPlan mine = buildPipelinedTransposePlan(g, pidx, ax, dir, in_halo, out_halo, in_pad, out_pad, inplace, nchunks,
                                                     false, elem_bytes, min_row_bytes);
    if (mine.steps.empty()) return mine;
    const int ci = mine.base.axes.comm;
    for (auto& st : mine.steps) st.push.clear();
    for (int j = 0; j < mine.base.comm_size; ++j) {
      auto pj = pidx;
      pj[ci] = j;
      const PipelinedPlan theirs =
          buildPipelinedTransposePlan(g, pj, ax, dir, in_halo, out_halo, in_pad, out_pad, inplace, nchunks, false, elem_bytes,
                                      min_row_bytes);
      for (size_t s = 0; s < theirs.steps.size() && s < mine.steps.size(); ++s)
        for (BoxDesc b : theirs.steps[s].push) {
          if (b.peer != mine.base.me) continue;
          b.peer = j;
          b.peer_world = mine.base.group_world[j];
          mine.steps[s].push.push_back(b);
        }
    }
    return mine;
  }
  PipelinedPlan pp;
  pp.base = buildTransposePlan(g, pidx, ax, dir, in_halo, out_halo, in_pad, out_pad, DstKind::STAGE, inplace);
  if (pp.base.noop || nchunks <= 1 || pp.base.push.empty()) return pp;
  int K = nchunks;
  const int a = pp.base.axes.a, b = pp.base.axes.b;
  const int P = pp.base.comm_size, me = pp.base.me;

  const Pencil pa = pencilInfo(g, pidx, a, nullptr, nullptr);
  const Pencil pa_h = pencilInfo(g, pidx, a, in_halo, in_pad);
  const Pencil pb = pencilInfo(g, pidx, b, nullptr, nullptr);
  const Pencil pb_h = pencilInfo(g, pidx, b, out_halo, out_pad);
  // Chunk axis. Default: the slowest axis of the source, so that chunks are ranges of source planes and the in-place
  // hazard analysis below works on address intervals. Better where it applies: the axis c that takes no part in the
  // transpose, when it is the FASTEST axis of both pencils (same memory order, same halo and padding along c). Both
  // pencils then consist of rows of one length, chunk k is the same column range of every row on both sides, and what
  // chunk k writes lies inside what chunk k has read (or outside the source pencil): every piece can be unpacked in the
  // step it arrives, with nothing piling up behind the last push. (With the default layout that is Y<->Z, chunked along
  // x; for X<->Y the planes along z coincide already.)
  // The row copy moves a row in warp-sized pieces of 2 KiB, and a chunk's rows are only as long as the chunk is wide:
  // measured on B200 (profiles/r2_n2_schedules.md), 2 KiB rows beat plane chunks by 5 %, 1.25 KiB rows LOSE 24 %. So
  // the chunk count is lowered to the largest one whose rows are whole multiples of min_row_bytes (2 KiB), and plane
  // chunks stay where there is none.
  const int c_axis = pp.base.axes.c;
  bool column_chunks = false;
  if (elem_bytes > 0 && pa.order == pb.order && pa.order[0] == c_axis && pa_h.halo[c_axis] == pb_h.halo[c_axis] &&
      pa_h.pad[c_axis] == pb_h.pad[c_axis]) {
    const int64_t row_bytes = static_cast<int64_t>(pa.shapeG()[c_axis]) * elem_bytes;
    for (int k = K; k >= 2 && !column_chunks; --k)
      if (row_bytes % (static_cast<int64_t>(k) * min_row_bytes) == 0 && pa.shapeG()[c_axis] % k == 0) {
        K = k;
        column_chunks = true;
      }
  }
  const int G = column_chunks ? c_axis : pa.order[2];
  pp.chunk_axis = G;
  const auto splits_a = getSplits(g.gdims_dist[a], P, g.gdims[a] - g.gdims_dist[a]);
  const auto splits_b = getSplits(g.gdims_dist[b], P, g.gdims[b] - g.gdims_dist[b]);
  const auto off_a = prefixOffsets(splits_a);
  const auto off_b = prefixOffsets(splits_b);
  const auto shape_a = pa.shapeG();
  const auto shape_b = pb.shapeG();
  const auto in_str = pa_h.strideG();
  pp.steps.resize(K);

  // ---- push: step k moves the k-th slice of EVERY box along G (slices are relative to the box, so each step is a
  // complete all-to-all of 1/K of the data: when G is the split axis `a` the boxes are consecutive plane ranges of my
  // pencil, and cutting the pencil instead of the boxes would send whole steps to a single peer)
  for (const BoxDesc& box : pp.base.push) {
    const int64_t len = box.ext[G];
    for (int k = 0; k < K; ++k) {
      auto [lo, hi] = chunkRange(len, K, k);
      if (hi <= lo) continue;
      BoxDesc sub = box;
      sub.ext[G] = hi - lo;
      sub.src_off += lo * box.sstr[G];
      sub.dst_off += lo * box.dstr[G];
      pp.steps[k].push.push_back(sub);
    }
  }

  // Source planes (along G, interior coordinates of my pencil) that no push has read after step `step`: one range per
  // box. G is the slowest axis of the source, so a plane range is one address interval of the input buffer.
  auto unreadIntervals = [&](int step) {
    std::vector<std::pair<int64_t, int64_t>> iv; // [first element, one past the last element) of the input buffer
    for (const BoxDesc& box : pp.base.push) {
      const int64_t start = (G == a) ? off_a[box.peer] : 0;
      const int64_t len = box.ext[G];
      const int64_t first = (step + 1 < K) ? chunkRange(len, K, step + 1).first : len;
      if (first >= len) continue;
      iv.push_back({(start + first + pa_h.halo[G]) * in_str[G], (start + len + pa_h.halo[G]) * in_str[G]});
      if (G != a) break; // every box spans the same planes
    }
    return iv;
  };

  // ---- unpack: what source j's slice k leaves in my workspace, and the first step at which its destination is free
  const auto dense = pb.strideG();
  const auto out_str = pb_h.strideG();
  for (int j = 0; j < P; ++j) {
    // extent along G of the box rank j sends to me, and where it starts in my (dense) destination pencil
    int64_t lenj, dst0;
    if (G == a) {
      lenj = splits_a[me]; // my share of a: the same slice of it arrives from every source
      dst0 = 0;
    } else if (G == b) {
      lenj = splits_b[j]; // rank j's own share of b
      dst0 = off_b[j];
    } else {
      lenj = shape_a[G];
      dst0 = 0;
    }
    for (int k = 0; k < K; ++k) {
      auto [u0, u1] = chunkRange(lenj, K, k);
      if (u1 <= u0) continue;
      // the piece in the coordinates of my (dense) destination pencil
      std::array<int64_t, 3> s0{}, ext{};
      for (int d = 0; d < 3; ++d) {
        s0[d] = 0;
        ext[d] = shape_b[d];
      }
      s0[b] = off_b[j];
      ext[b] = splits_b[j];
      s0[G] = dst0 + u0;
      ext[G] = u1 - u0;
      BoxDesc piece;
      piece.peer = me;
      piece.peer_world = pp.base.group_world[me];
      piece.ext = ext;
      piece.sstr = dense;
      piece.dstr = out_str;
      piece.src_off = dot3(s0, dense);
      std::array<int64_t, 3> d0{s0[0] + pb_h.halo[0], s0[1] + pb_h.halo[1], s0[2] + pb_h.halo[2]};
      piece.dst_off = dot3(d0, out_str);
      if (piece.count() == 0) continue;

      // first step at which everything `pc` writes has been read by the pushes (in place only)
      auto firstFreeStep = [&](const BoxDesc& pc) {
        const int64_t first = pc.dst_off;
        const int64_t last = pc.dst_off + (pc.ext[0] - 1) * out_str[0] + (pc.ext[1] - 1) * out_str[1] + (pc.ext[2] - 1) * out_str[2];
        int st = k;
        for (; st < K - 1; ++st) {
          bool clear = true;
          for (auto& iv : unreadIntervals(st))
            if (first < iv.second && last >= iv.first) clear = false;
          if (clear) break;
        }
        return st;
      };
      int step = k;
      if (inplace && !column_chunks) {
        step = firstFreeStep(piece);
        // A piece that has to wait is cut along the slowest axis of the destination layout and every part waits only
        // for the source planes IT overwrites: a piece spans P source slices' worth of memory when G is the split
        // axis, so uncut it would wait for the last of them (P/K of the pencil left for after the last push; 1/K cut).
        const int D = pb.order[2];
        const int64_t parts = std::min<int64_t>(piece.ext[D], std::min(std::max(P, K), 16));
        if (step > k && parts > 1) {
          for (int64_t q = 0; q < parts; ++q) {
            auto [q0, q1] = chunkRange(piece.ext[D], static_cast<int>(parts), static_cast<int>(q));
            if (q1 <= q0) continue;
            BoxDesc part = piece;
            part.ext[D] = q1 - q0;
            part.src_off += q0 * dense[D];
            part.dst_off += q0 * out_str[D];
            pp.steps[firstFreeStep(part)].unpack.push_back(part);
          }
          continue;
        }
      }
      pp.steps[step].unpack.push_back(piece);
    }
  }
  return pp;
}

HaloPlan buildHaloPlan(const GridGeom& g, const std::array<int, 2>& pidx, int ax, int dim, const int32_t halo[3],
                       const bool periods[3], const int32_t pad[3], DstKind kind) {
  HaloPlan plan;
  const Pencil ph = pencilInfo(g, pidx, ax, halo, nullptr);
  const Pencil php = pencilInfo(g, pidx, ax, halo, pad);
  if (hasEmptyPencils(g, ax)) THROW_NOT_SUPPORTED("halo operations on configurations with empty pencils not supported");

  const bool periodic = periods ? periods[dim] : false;
  const int hw = ph.halo[dim];
  const int my_world = rankOfPidx(g, pidx);

  if (dim == ax) {
    plan.comm_size = 1;
    plan.me = 0;
    plan.group_world = {my_world};
    plan.neighbor = periodic ? std::array<int, 2>{0, 0} : std::array<int, 2>{-1, -1};
  } else {
    plan.comm = haloCommAxis(ax, dim);
    const int P = g.pdims[plan.comm];
    plan.comm_size = P;
    plan.me = pidx[plan.comm];
    for (int i = 0; i < P; ++i) {
      auto pp = pidx;
      pp[plan.comm] = i;
      plan.group_world.push_back(rankOfPidx(g, pp));
    }
    int l = plan.me - 1, r = plan.me + 1;
    if (periodic) {
      l = (l + P) % P;
      r = r % P;
    } else {
      if (l < 0) l = -1;
      if (r >= P) r = -1;
    }
    plan.neighbor = {l, r};
  }

  if (hw == 0) { // reference halo.h:70-71
    plan.nothing = true;
    return plan;
  }
  if (plan.neighbor[0] == -1 && plan.neighbor[1] == -1) {
    plan.nothing = true;
    return plan;
  }
  const bool self_only = (plan.neighbor[0] == plan.me && plan.neighbor[1] == plan.me);
  if (!self_only) {
    // the face must come from the nearest neighbour alone (reference halo.h:120-145)
    const auto splits = getSplits(g.gdims_dist[dim], plan.comm_size, g.gdims[dim] - g.gdims_dist[dim]);
    for (int nb : plan.neighbor) {
      if (nb < 0) continue;
      if (hw > splits[nb] || hw > splits[plan.me])
        THROW_INVALID_USAGE(
            "halo includes ranks other than nearest neighbor processes, this is not currently supported.");
    }
  }

  const auto shape_h = ph.shapeG();   // halo-inclusive, padding-exclusive
  const auto shape_hp = php.shapeG(); // what the buffer really is
  const auto sstr = php.strideG();
  std::array<int64_t, 3> ext{};
  for (int k = 0; k < 3; ++k) ext[k] = (k == dim) ? hw : shape_h[k];
  plan.face_elems = ext[0] * ext[1] * ext[2];
  const int64_t slot = alignCount(plan.face_elems);

  // dense strides of a face in the pencil's memory order
  std::array<int64_t, 3> dense{};
  {
    int64_t acc = 1;
    for (int k = 0; k < 3; ++k) {
      dense[php.order[k]] = acc;
      acc *= ext[php.order[k]];
    }
  }

  for (int side = 0; side < 2; ++side) {
    const int nb = plan.neighbor[side];
    if (nb < 0) continue;
    auto pp = pidx;
    if (dim != ax) pp[plan.comm] = nb;
    const Pencil q = pencilInfo(g, pp, ax, halo, pad);
    const auto qshape = q.shapeG();

    BoxDesc bx;
    bx.peer = nb;
    bx.peer_world = plan.group_world[nb];
    bx.ext = ext;
    bx.sstr = sstr;
    std::array<int64_t, 3> s0{}, d0{};
    // side 0: my first interior layers go to the left neighbour's right halo;
    // side 1: my last interior layers go to the right neighbour's left halo
    s0[dim] = (side == 0) ? hw : shape_hp[dim] - 2 * hw - php.pad[dim];
    bx.src_off = dot3(s0, sstr);
    if (kind == DstKind::FINAL) {
      bx.dstr = q.strideG();
      d0[dim] = (side == 0) ? qshape[dim] - hw - q.pad[dim] : 0;
      bx.dst_off = dot3(d0, bx.dstr);
    } else {
      bx.dstr = dense;
      bx.dst_off = (side == 0) ? slot : 0; // receiver slot 0 = its left halo, slot 1 = its right halo
    }
    plan.push.push_back(bx);
  }

  if (kind == DstKind::STAGE) {
    for (int side = 0; side < 2; ++side) {
      if (plan.neighbor[side] < 0) continue;
      BoxDesc u;
      u.peer = plan.me;
      u.peer_world = my_world;
      u.ext = ext;
      u.sstr = dense;
      u.src_off = (side == 0) ? 0 : slot;
      u.dstr = sstr;
      std::array<int64_t, 3> d0{};
      d0[dim] = (side == 0) ? 0 : shape_hp[dim] - hw - php.pad[dim];
      u.dst_off = dot3(d0, sstr);
      plan.unpack.push_back(u);
    }
  }
  return plan;
}

CanonBox canonicalize(const BoxDesc& b, bool merge) {
  struct Ax {
    int64_t n, ss, ds;
  };
  std::vector<Ax> axes;
  for (int k = 0; k < 3; ++k)
    if (b.ext[k] > 1) axes.push_back({b.ext[k], b.sstr[k], b.dstr[k]});
  std::sort(axes.begin(), axes.end(), [](const Ax& x, const Ax& y) { return x.ss != y.ss ? x.ss < y.ss : x.ds < y.ds; });
  if (merge) {
    for (size_t k = 0; k + 1 < axes.size();) {
      if (axes[k + 1].ss == axes[k].n * axes[k].ss && axes[k + 1].ds == axes[k].n * axes[k].ds) {
        axes[k].n *= axes[k + 1].n;
        axes.erase(axes.begin() + k + 1);
      } else {
        ++k;
      }
    }
  }
  // a box whose unit-stride axis has extent 1 still needs a (trivial) row axis
  const bool has_row = !axes.empty() && axes[0].ss == 1 && axes[0].ds == 1;
  bool transposable = false;
  if (!axes.empty() && axes[0].ss == 1)
    for (size_t k = 1; k < axes.size(); ++k)
      if (axes[k].ds == 1) transposable = true;
  if (!has_row && !transposable && axes.size() < 3) axes.insert(axes.begin(), Ax{1, 1, 1});
  if (b.count() == 0) axes.assign(1, Ax{0, 1, 1});
  if (axes.empty()) axes.push_back({1, 1, 1});

  CanonBox c;
  c.nd = static_cast<int>(axes.size());
  for (int k = 0; k < 3; ++k) {
    if (k < c.nd) {
      c.n[k] = axes[k].n;
      c.ss[k] = axes[k].ss;
      c.ds[k] = axes[k].ds;
    } else {
      c.n[k] = 1;
      c.ss[k] = 0;
      c.ds[k] = 0;
    }
  }
  return c;
}

} // namespace cdb
