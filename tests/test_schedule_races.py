"""A happens-before checker for the engine's schedules (CPU only).

The copy kernels of different ranks and streams run concurrently; what keeps them from racing is (a) stream order,
(b) events between the caller's stream and the side stream of the chunked schedule, and (c) the in-kernel handshake:
  entry : no access of a launch begins before every member of its communicator has STARTED the same launch
          (a member's launch starts after everything earlier on its stream);
  exit  : a launch does not COMPLETE before every member has finished its accesses;
  step  : (fused staged schedule, kernels.cu rowCopyPhasedKernel) inside ONE launch per rank, the unpack of chunk s
          begins only after every member has finished pushing chunks 0..s; pushes of different chunks and unpacks of
          different chunks are otherwise unordered (CTAs drift), and there is NO exit handshake.
This test restates the launch structure of engine.cc (runTranspose: direct / staged / chunked, sender- and
receiver-driven) as a graph of such orderings, takes the boxes of every launch from the planner entry points of
libcudecomp.so, and requires a happens-before path between any two accesses of different launches that touch the same
cell when at least one of them writes. Chains of operations (X->Y->Z->Y->X, twice, alternating buffers like a real
caller) are checked as a whole, so hazards ACROSS operations are covered too: a peer still reading a buffer that the next
operation overwrites would show up here. The checker itself is validated on schedules that are known to race
(direct stores into a buffer that is transposed in place; an unpack that does not wait for its data).
"""
import itertools

import numpy as np
import pytest

from cudecomp_b200 import capi as cd
from oracle import oracle as orc
from tests.test_planner_properties import OPS, make_config, make_oracle

CHAIN = ["XY", "YZ", "ZY", "YX", "XY", "YZ", "ZY", "YX"]


def cells(box, which):
    idx = np.indices(box["extent"], dtype=np.int64).reshape(3, -1)
    off, strides = (box["src_offset"], box["src_stride"]) if which == "src" else (box["dst_offset"], box["dst_stride"])
    return np.unique(off + sum(idx[k] * strides[k] for k in range(3)))


class Graph:
    """Nodes are moments (start / first access / last access / completion of a launch); edges are 'not later than'."""

    def __init__(self):
        self.succ = []
        self.launches = []

    def node(self):
        self.succ.append(set())
        return len(self.succ) - 1

    def edge(self, a, b):
        self.succ[a].add(b)

    def launch(self, rank, stream, name, after):
        """A kernel: start <= first access <= last access <= completion; `after` are nodes that precede its start."""
        L = dict(rank=rank, stream=stream, name=name, start=self.node(), begin=self.node(), end=self.node(),
                 done=self.node(), reads=[], writes=[])
        self.edge(L["start"], L["begin"])
        self.edge(L["begin"], L["end"])
        self.edge(L["end"], L["done"])
        for a in after:
            self.edge(a, L["start"])
        self.launches.append(L)
        return L

    def handshake(self, group, entry=True, exit_=True):
        for A in group:
            for B in group:
                if A is B:
                    continue
                if entry:
                    self.edge(B["start"], A["begin"])
                if exit_:
                    self.edge(B["end"], A["done"])

    def reachability(self):
        n = len(self.succ)
        order, seen = [], [False] * n
        for root in range(n):  # iterative DFS post-order = reverse topological order
            if seen[root]:
                continue
            stack = [(root, iter(self.succ[root]))]
            seen[root] = True
            while stack:
                v, it = stack[-1]
                for w in it:
                    if not seen[w]:
                        seen[w] = True
                        stack.append((w, iter(self.succ[w])))
                        break
                else:
                    order.append(v)
                    stack.pop()
        reach = [0] * n
        for v in order:
            bits = 1 << v
            for w in self.succ[v]:
                bits |= reach[w]
            reach[v] = bits
        return reach

    def races(self):
        reach = self.reachability()
        found = []
        for L1, L2 in itertools.combinations(self.launches, 2):
            ordered = (reach[L1["end"]] >> L2["begin"]) & 1 or (reach[L2["end"]] >> L1["begin"]) & 1
            if ordered:
                continue
            for (k1, acc1), (k2, acc2) in itertools.product((("r", L1["reads"]), ("w", L1["writes"])),
                                                            (("r", L2["reads"]), ("w", L2["writes"]))):
                if k1 == "r" and k2 == "r":
                    continue
                for (b1, c1), (b2, c2) in itertools.product(acc1, acc2):
                    if b1 == b2 and np.intersect1d(c1, c2, assume_unique=True).size:
                        found.append((L1["name"], L2["name"], b1))
        return found


def build_chain(d, mode, inplace, K=4, broken=None):
    """Graph of the whole CHAIN on every rank. mode: direct | pull | staged | chunked. Buffers are (rank, name) with
    names a, b (the caller's two pencils; in place only a) and work."""
    cfg, o = make_config(d), make_oracle(d)
    n = o.nranks
    g = Graph()
    last_main = [None] * n   # completion of the last launch on each rank's caller stream
    cur = ["a"] * n
    for opi, op in enumerate(CHAIN):
        ax, direction = OPS[op]
        src_name = cur[0]
        dst_name = src_name if inplace else ("b" if src_name == "a" else "a")
        # communicator size of this operation (all groups have the same size)
        P = len(cd.plan_transpose_boxes(cfg, 0, ax, direction))
        eff = mode
        if P == 1:
            # engine.cc: one rank per communicator -> local path: out of place one direct copy, in place through `work`
            eff = "staged" if inplace else "direct"
        if eff == "fused":
            fused_launches(g, cfg, n, op, opi, ax, direction, inplace, K, src_name, dst_name, last_main, broken)
            cur = [dst_name] * n
            continue
        plans = {}
        for r in range(n):
            if eff in ("chunked", "pull_chunked"):
                plans[r] = cd.plan_pipelined_transpose_boxes(cfg, r, ax, direction, None, None, None, None,
                                                             int(inplace) + (2 if eff == "pull_chunked" else 0), K)
            else:
                plans[r] = cd.plan_transpose_boxes(cfg, r, ax, direction, None, None, None, None,
                                                   {"direct": 0, "staged": 1, "pull": 2, "pull_staged": 3}[eff])
        groups = {}
        for r in range(n):
            members = tuple(sorted({b["peer_rank"] for b in plans[r] if not b["is_unpack"]} | {r}))
            groups.setdefault(members, []).append(r)
        steps = K if eff in ("chunked", "pull_chunked") else 1
        side_last = [None] * n
        for s in range(steps):
            step_launch = {}
            for r in range(n):
                push = [b for b in plans[r] if not b["is_unpack"] and (steps == 1 or b["step"] == s)]
                L = g.launch(r, "main", "%s#%d push%d r%d" % (op, opi, s, r), [last_main[r]] if last_main[r] is not None else [])
                for b in push:
                    if eff in ("pull", "pull_staged", "pull_chunked"):
                        L["reads"].append(((b["peer_rank"], src_name), cells(b, "src")))
                        L["writes"].append(((r, dst_name if eff == "pull" else "work"), cells(b, "dst")))
                    else:
                        target = "work" if eff in ("staged", "chunked") else dst_name
                        L["reads"].append(((r, src_name), cells(b, "src")))
                        L["writes"].append(((b["peer_rank"], target), cells(b, "dst")))
                step_launch[r] = L
                last_main[r] = L["done"]
            for members in groups.values():
                if len(members) > 1:
                    g.handshake([step_launch[r] for r in members], entry=(s == 0), exit_=True)
            for r in range(n):
                unpack = [b for b in plans[r] if b["is_unpack"] and (steps == 1 or b["step"] == s)]
                if not unpack:
                    continue
                if steps == 1:
                    U = g.launch(r, "main", "%s#%d unpack r%d" % (op, opi, r), [last_main[r]])
                    last_main[r] = U["done"]
                else:
                    after = [step_launch[r]["done"]] if broken != "unpack_does_not_wait" else []
                    if side_last[r] is not None:
                        after.append(side_last[r])
                    U = g.launch(r, "side", "%s#%d unpack%d r%d" % (op, opi, s, r), after)
                    side_last[r] = U["done"]
                for b in unpack:
                    U["reads"].append(((r, "work"), cells(b, "src")))
                    U["writes"].append(((r, dst_name), cells(b, "dst")))
        for r in range(n):
            if side_last[r] is not None:  # the caller's stream rejoins the side stream at the end of the call
                J = g.launch(r, "main", "%s#%d join r%d" % (op, opi, r), [last_main[r], side_last[r]])
                last_main[r] = J["done"]
        cur = [dst_name] * n
    return g


def fused_launches(g, cfg, n, op, opi, ax, direction, inplace, K, src_name, dst_name, last_main, broken):
    """engine.cc runFusedStaged: one phased launch per rank. Column chunks where the planner uses them (element size 8,
    no row-length floor on these small grids)."""
    flags = int(inplace) + (8 << 8) + 4
    plans = {r: cd.plan_pipelined_transpose_boxes(cfg, r, ax, direction, None, None, None, None, flags, K) for r in range(n)}
    if not any(plans.values()):  # chunking does not apply: the plain staged plan as ONE step
        plans = {r: [dict(b, step=0) for b in cd.plan_transpose_boxes(cfg, r, ax, direction, None, None, None, None, 1)]
                 for r in range(n)}
    steps = 1 + max(b["step"] for r in range(n) for b in plans[r])
    groups = {}
    for r in range(n):
        members = tuple(sorted({b["peer_rank"] for b in plans[r] if not b["is_unpack"]} | {r}))
        groups[r] = members
    kern, push, unpack = {}, {}, {}
    for r in range(n):
        kern[r] = g.launch(r, "main", "%s#%d fused r%d" % (op, opi, r), [last_main[r]] if last_main[r] is not None else [])
        for s in range(steps):
            P = g.launch(r, "main", "%s#%d push%d r%d" % (op, opi, s, r), [kern[r]["start"]])
            g.edge(P["done"], kern[r]["done"])
            for b in plans[r]:
                if not b["is_unpack"] and b["step"] == s:
                    P["reads"].append(((r, src_name), cells(b, "src")))
                    P["writes"].append(((b["peer_rank"], "work"), cells(b, "dst")))
            push[r, s] = P
            U = g.launch(r, "main", "%s#%d unpack%d r%d" % (op, opi, s, r), [kern[r]["start"]])
            g.edge(U["done"], kern[r]["done"])
            for b in plans[r]:
                if b["is_unpack"] and b["step"] == s:
                    U["reads"].append(((r, "work"), cells(b, "src")))
                    U["writes"].append(((r, dst_name), cells(b, "dst")))
            unpack[r, s] = U
        last_main[r] = kern[r]["done"]
    for r in range(n):
        for m in groups[r]:
            for s in range(steps):
                if m != r and broken != "fused_no_entry":
                    # entry handshake: nothing of mine touches memory before every member has started its launch
                    g.edge(kern[m]["start"], push[r, s]["begin"])
                    g.edge(kern[m]["start"], unpack[r, s]["begin"])
                # step flags + local counter: chunk s is unpacked after every member has pushed chunks 0..s
                for t in range(0, s + 1):
                    if broken == "fused_unpack_early" and t == s:
                        continue
                    g.edge(push[m, t]["end"], unpack[r, s]["begin"])


def decomposition(gdims, pdims, axis_contiguous=False):
    return dict(gdims=gdims, pdims=pdims, axis_contiguous=[axis_contiguous] * 3, mem_order=None, gdims_dist=None,
                col_major=False, halos={str(a): [0, 0, 0] for a in range(3)}, pads={str(a): [0, 0, 0] for a in range(3)})


GRIDS = [([8, 6, 10], [2, 2], False), ([7, 9, 8], [2, 2], True), ([8, 8, 8], [1, 4], False), ([9, 8, 6], [4, 1], False),
         ([6, 6, 6], [2, 1], True)]


@pytest.mark.parametrize("gdims,pdims,ac", GRIDS, ids=["%dx%d%s" % (p[0], p[1], "_ac" if a else "") for _, p, a in GRIDS])
@pytest.mark.parametrize("mode,inplace", [("direct", False), ("pull", False), ("staged", True), ("staged", False),
                                          ("pull_staged", True), ("pull_staged", False), ("chunked", True),
                                          ("chunked", False), ("pull_chunked", True), ("pull_chunked", False),
                                          ("fused", True), ("fused", False)])
def test_schedules_are_race_free(gdims, pdims, ac, mode, inplace):
    g = build_chain(decomposition(gdims, pdims, ac), mode, inplace, K=3)
    assert g.races() == []


def test_the_checker_sees_known_races():
    d = decomposition([8, 6, 10], [2, 2])
    # peers storing straight into a pencil that its owner is still reading (why in-place calls are staged)
    assert build_chain(d, "direct", True).races()
    # receiver-driven in place: a peer reads my input while I overwrite it with my output
    assert build_chain(d, "pull", True).races()
    # an unpack that does not wait for the push that delivers its data
    assert build_chain(d, "chunked", True, K=3, broken="unpack_does_not_wait").races()
    # fused schedule: an unpack that only waits for the chunk BEFORE its own; no entry handshake (a peer of the next
    # operation pushes into a workspace that is still being unpacked -- the exit handshake the fused launch omits is not
    # what protects it, the next launch's entry handshake is)
    assert build_chain(d, "fused", True, K=3, broken="fused_unpack_early").races()
    assert build_chain(d, "fused", True, K=3, broken="fused_no_entry").races()
