"""Expert-parallel MoEMultiBranchFFN over NVLink peer memory (BASELINE.json configs[3]).

The reference keeps all experts of a SwitchMoELayer in one nn.ModuleList and loops over them in Python
(models/switch_moe.py:19-25,97-109; models/multi_branch.py:52-61); it has no expert parallelism.  Here
expert e of every branch lives on rank e // (E // R) of one NVSwitch node, tokens stay sharded by sequence,
and the "all-to-all" is not a separate collective: the dispatch kernel writes each routed row directly into
the expert-sorted buffer of the owning GPU (peer-mapped pointer), the combine kernel reads the expert
outputs back from the owners (csrc/ep.cu).  Cross-rank ordering uses a flag barrier in peer memory.

Two ways to obtain the R cooperating instances:
  * `ExpertParallelFFN.create_distributed(...)`: one process per GPU (torchrun); buffers are exchanged as
    CUDA IPC handles through `torch.distributed.all_gather_object` on the given process group.
  * `ExpertParallelFFN.create_emulated(R, ...)`: R virtual ranks inside one process on one GPU (all "peer"
    pointers are local).  `run_emulated` executes the phases rank by rank; used by the single-GPU tests and
    to check bit-equality against the local (non-EP) kernels.
"""
import ctypes as C
from typing import List, Optional

import torch

from . import _lib, ops
from ._lib import ACT_GELU, MDM_BF16, MDM_F32, EpPeers, MdmError


def _round_up(a, b):
    return (a + b - 1) // b * b


def owner_of(e: int, E: int, R: int) -> int:
    return e // (E // R)


def plan_segments(cnt, NB: int, E: int, R: int):
    """Host mirror of csrc/ep.cu:ep_scan_kernel (documentation + CPU tests): from the [R][NB*E] table of
    per-rank group counts, returns (dest_base[R][GT], seg_off[GT], seg_rows[GT], tiles_per_rank[R]);
    dest_base[s][g] is the first row, inside owner(g)'s buffer, of the rows rank s routes to group g."""
    EPR, GT = E // R, NB * E
    own = [owner_of(g % E, E, R) for g in range(GT)]
    lg = [(g // E) * EPR + (g % E) % EPR for g in range(GT)]
    rows = [sum(cnt[s][g] for s in range(R)) for g in range(GT)]
    padded = [_round_up(r, 128) for r in rows]
    seg_off = [sum(padded[j] for j in range(GT) if own[j] == own[g] and lg[j] < lg[g]) for g in range(GT)]
    dest = [[seg_off[g] + sum(cnt[t][g] for t in range(s)) for g in range(GT)] for s in range(R)]
    tiles = [sum(padded[g] for g in range(GT) if own[g] == r) // 128 for r in range(R)]
    return dest, seg_off, rows, tiles


class ExpertParallelFFN:
    def __init__(self, D, F, E, NB, R, me, n_tokens, dtype, device):
        if E % R:
            raise MdmError("the expert count (%d) must be divisible by the number of ranks (%d)" % (E, R))
        if R > _lib.EP_MAX_RANKS:
            raise MdmError("at most %d ranks" % _lib.EP_MAX_RANKS)
        self.D, self.F, self.E, self.NB, self.R, self.me = D, F, E, NB, R, me
        self.EPR, self.GT, self.GL = E // R, NB * E, NB * (E // R)
        self.N = n_tokens
        self.dtype, self.device = dtype, device
        self.dt = MDM_BF16 if dtype == torch.bfloat16 else MDM_F32
        # worst case: every token of every rank sends min(K, EPR) rows per branch to this rank
        self.cap = _round_up(R * n_tokens * NB * min(2, self.EPR), 128) + self.GL * 128
        i32, f32 = torch.int32, torch.float32
        z = lambda *s, dtype=f32: torch.zeros(*s, dtype=dtype, device=device)
        # peer-visible buffers
        self.xp = z(self.cap, D, dtype=dtype)
        self.rowscale = z(self.cap)
        self.yp = z(self.cap, D, dtype=dtype)
        self.cnt = z(R, self.GT, dtype=i32)
        self.flags = z(R, dtype=i32)
        # local buffers
        N, nblk = n_tokens, (n_tokens + 127) // 128
        self.hp = z(self.cap, F, dtype=dtype)
        self.idx = z(N, NB, 2, dtype=i32)
        self.vals = z(N, NB, 2)
        self.stats = z(N, 2)
        self.hist = z(nblk, 2, self.GT, dtype=i32)
        self.imp = z(nblk, self.GT)
        self.blk_base = z(nblk, self.GT, dtype=i32)
        self.dest_base = z(self.GT, dtype=i32)
        self.max_tiles = self.cap // 128
        self.t_up = z(self.max_tiles, 4, dtype=i32)
        self.t_dn = z(self.max_tiles, 4, dtype=i32)
        self.ntile = z(1, dtype=i32)
        self.overflow = z(1, dtype=i32)
        self.err = z(1, dtype=i32)
        self.perm = z(N, NB * 2, dtype=i32)
        self.usage = z(self.GT)
        self.importance = z(self.GT)
        self.epoch_ctr = z(1, dtype=i32)      # device-side barrier epoch (graph-replay safe)
        self.peers = None
        self._opened = []
        self.group = None

    # ------------------------------------------------------------------ construction
    def shard_weights(self, ln_w, ln_b, gate_w, gate_b, w1, b1, w2, b2):
        """Full (all-expert) tensors in the single-GPU packed layout (MotionTransformer._pack): w1 [GT*F, D],
        b1 [GT*F], w2 [GT*D, F], b2 [GT*D], group g = branch * E + expert.  Returns the dict of tensors this
        rank needs: gate / LayerNorm of every expert (routing is local), FFN weights of the owned experts only."""
        D, F, E, EPR = self.D, self.F, self.E, self.EPR
        own = [g for g in range(self.GT) if owner_of(g % E, E, self.R) == self.me]
        own.sort(key=lambda g: (g // E) * EPR + (g % E) % EPR)
        sel = lambda t, width: torch.cat([t[g * width:(g + 1) * width] for g in own]).contiguous()
        return dict(ln_w=ln_w.float().contiguous(), ln_b=ln_b.float().contiguous(),
                    gate_w=gate_w.float().contiguous(), gate_b=gate_b.float().contiguous(),
                    w1=sel(w1, F).to(self.dtype), b1=sel(b1, F).float(),
                    w2=sel(w2, D).to(self.dtype), b2=sel(b2, D).float())

    def use_weights(self, wd):
        """Select the (already sharded) weights of the MoE layer the next calls belong to."""
        for k, t in wd.items():
            setattr(self, k, t)

    def set_weights(self, ln_w, ln_b, gate_w, gate_b, w1, b1, w2, b2):
        self.use_weights(self.shard_weights(ln_w, ln_b, gate_w, gate_b, w1, b1, w2, b2))

    def _fill_peers(self, ptrs):
        """ptrs[p] = dict(xp=, rowscale=, yp=, cnt=, flags=) of device addresses valid on this GPU."""
        pe = EpPeers()
        for p in range(self.R):
            pe.xp[p], pe.rowscale[p], pe.yp[p] = ptrs[p]["xp"], ptrs[p]["rowscale"], ptrs[p]["yp"]
            pe.cnt[p], pe.flags[p] = ptrs[p]["cnt"], ptrs[p]["flags"]
        self.peers = pe

    def _local_ptrs(self):
        return {k: getattr(self, k).data_ptr() for k in ("xp", "rowscale", "yp", "cnt", "flags")}

    @classmethod
    def create_emulated(cls, R, D, F, E, NB, n_tokens, dtype, device) -> List["ExpertParallelFFN"]:
        inst = [cls(D, F, E, NB, R, r, n_tokens, dtype, device) for r in range(R)]
        ptrs = [i._local_ptrs() for i in inst]
        for i in inst:
            i._fill_peers(ptrs)
        return inst

    @classmethod
    def create_distributed(cls, D, F, E, NB, n_tokens, dtype, device, group=None) -> "ExpertParallelFFN":
        import torch.distributed as dist
        R, me = dist.get_world_size(group), dist.get_rank(group)
        self = cls(D, F, E, NB, R, me, n_tokens, dtype, device)
        self.group = group
        lib = _lib.load()
        mine = {}
        for k in ("xp", "rowscale", "yp", "cnt", "flags"):
            h = (C.c_ubyte * 64)()
            off = C.c_long(0)
            _lib.check(lib.mdm_ipc_get_handle(getattr(self, k).data_ptr(), h, C.byref(off)), "mdm_ipc_get_handle")
            mine[k] = (bytes(h), off.value)
        mine["_shape"] = (int(n_tokens), int(self.cap), D, F, E, NB, str(dtype))
        everyone = [None] * R
        dist.all_gather_object(everyone, mine, group=group)
        # every rank must use the same token count / buffer capacity: dispatch bounds rows by the SENDER's cap while writing
        # into the OWNER's buffers, so a rank with a larger shard would write out of bounds on a peer
        shapes = [e.pop("_shape") for e in everyone]
        if any(sh != shapes[0] for sh in shapes):
            raise MdmError("expert parallelism needs identical shapes on every rank (n_tokens, cap, D, F, E, NB, dtype); "
                           "got %s - shard the batch evenly (B %% world == 0)" % (shapes,))
        ptrs, opened = [], {}
        for p in range(R):
            if p == me:
                ptrs.append(self._local_ptrs())
                continue
            d = {}
            for k, (h, off) in everyone[p].items():
                if h not in opened:      # several tensors may share one cudaMalloc block of the caching allocator
                    base = C.c_void_p()
                    buf = (C.c_ubyte * 64).from_buffer_copy(h)
                    _lib.check(lib.mdm_ipc_open_handle(buf, C.byref(base)), "mdm_ipc_open_handle")
                    opened[h] = base.value
                    self._opened.append(base.value)
                d[k] = opened[h] + off
            ptrs.append(d)
        self._fill_peers(ptrs)
        # The zero-fills of this rank's peer-visible buffers (cnt, flags, xp, ...) are asynchronous on its stream; a peer that
        # runs ahead writes its count row / barrier flag into them through NVLink as soon as its first MoE call starts.  Make
        # sure every rank's buffers are initialised ON THE DEVICE before any rank leaves this constructor (a host barrier
        # alone does not order the GPUs: found with two ranks time-sliced on one GPU, where the first forward occasionally
        # routed with a zeroed count table).
        torch.cuda.synchronize(device)
        dist.barrier(group=group)
        return self

    def close(self):
        lib = _lib.load()
        for b in self._opened:
            lib.mdm_ipc_close_handle(b)
        self._opened = []

    # ------------------------------------------------------------------ phases of one MoE call
    def phase_gate(self, x, usage=None, importance=None):
        """x [N, D] fp32 (the residual stream before the MoE FFN); usage / importance: the [NB*E] counter
        rows of this layer (default: the instance's own)."""
        ops._c(x, usage, importance)
        usage = self.usage if usage is None else usage
        importance = self.importance if importance is None else importance
        N, D, NB, E = self.N, self.D, self.NB, self.E
        ops.moe_gate(x, N, D, NB, E, self.ln_w, self.ln_b, self.gate_w, self.gate_b, self.idx, self.vals, self.stats,
                     self.hist, self.imp)
        lib = _lib.load()
        _lib.check(lib.mdm_ep_counts(self.hist.data_ptr(), self.imp.data_ptr(), N, NB, E, 2, self.R, self.me,
                                     C.byref(self.peers), self.blk_base.data_ptr(), usage.data_ptr(),
                                     importance.data_ptr(), ops._stream()), "mdm_ep_counts")
        self._x = x

    def phase_dispatch(self):
        lib, N, D, NB, E = _lib.load(), self.N, self.D, self.NB, self.E
        _lib.check(lib.mdm_ep_scan(self.cnt.data_ptr(), NB, E, 2, self.R, self.me, self.F, D, self.cap,
                                   self.dest_base.data_ptr(), self.t_up.data_ptr(), self.t_dn.data_ptr(),
                                   self.ntile.data_ptr(), self.overflow.data_ptr(), ops._stream()), "mdm_ep_scan")
        _lib.check(lib.mdm_ep_dispatch(self._x.data_ptr(), N, D, NB, E, 2, self.R, self.me, self.cap,
                                       self.ln_w.data_ptr(), self.ln_b.data_ptr(), self.idx.data_ptr(),
                                       self.vals.data_ptr(), self.stats.data_ptr(), self.blk_base.data_ptr(),
                                       self.dest_base.data_ptr(), C.byref(self.peers), self.dt,
                                       self.perm.data_ptr(), ops._stream()), "mdm_ep_dispatch")

    def phase_experts(self):
        kw = dict(num_tiles=self.max_tiles, num_tiles_dev=self.ntile, M=self.cap, a_rows=self.cap)
        if self.dtype == torch.bfloat16:
            ops.gemm(self.xp, self.w1, self.b1, act=ACT_GELU, out_a=self.hp, N=self.F, tiles=self.t_up,
                     w_rows=self.GL * self.F, **kw)
            ops.gemm(self.hp, self.w2, self.b2, out_a=self.yp, N=self.D, rowscale=self.rowscale, tiles=self.t_dn,
                     w_rows=self.GL * self.D, **kw)
        else:
            ops.gemm(self.xp, self.w1, self.b1, act=ACT_GELU, out_f32=self.hp, N=self.F, tiles=self.t_up,
                     w_rows=self.GL * self.F, **kw)
            ops.gemm(self.hp, self.w2, self.b2, out_f32=self.yp, N=self.D, rowscale=self.rowscale, tiles=self.t_dn,
                     w_rows=self.GL * self.D, **kw)

    def phase_combine(self, s_norm_w, s_norm_b, film, rows_per_seq, out):
        ops._c(s_norm_w, s_norm_b, film, out)
        _lib.check(_lib.load().mdm_ep_combine_film(C.byref(self.peers), self.dt, self.perm.data_ptr(), self.N, self.D,
                                                   self.NB * 2, self.cap, s_norm_w.data_ptr(), s_norm_b.data_ptr(),
                                                   film.data_ptr(), rows_per_seq, out.data_ptr(), ops._stream()),
                   "mdm_ep_combine_film")

    def barrier(self):
        _lib.check(_lib.load().mdm_ep_barrier(C.byref(self.peers), self.R, self.me, self.epoch_ctr.data_ptr(),
                                              self.err.data_ptr(), ops._stream()), "mdm_ep_barrier")

    def forward(self, x, s_norm_w, s_norm_b, film, rows_per_seq, out, usage=None, importance=None):
        """One MoEMultiBranchFFN call up to (not including) its output Linear, in stream order; every rank
        must call it the same number of times (the barriers count epochs).  CUDA-graph capturable."""
        self.phase_gate(x, usage, importance)
        self.barrier()
        self.phase_dispatch()
        self.barrier()
        self.phase_experts()
        self.barrier()
        self.phase_combine(s_norm_w, s_norm_b, film, rows_per_seq, out)
        return out

    def check_health(self):
        """Host-side check (synchronises): no barrier time-out and no buffer overflow so far."""
        if int(self.err.item()):
            raise MdmError("expert-parallel barrier timed out: a peer rank did not arrive")
        if int(self.overflow.item()):
            raise MdmError("expert-parallel receive buffer overflow (cap=%d rows)" % self.cap)


def run_emulated(inst: List[ExpertParallelFFN], xs, s_norm_w, s_norm_b, films, rows_per_seq, outs):
    """All R virtual ranks of one process, phase by phase (the sequential order replaces the barriers)."""
    for i, x in zip(inst, xs):
        i.phase_gate(x)
    for i in inst:
        i.phase_dispatch()
    for i in inst:
        i.phase_experts()
    for i, f, o in zip(inst, films, outs):
        i.phase_combine(s_norm_w, s_norm_b, f, rows_per_seq, o)
    return outs
