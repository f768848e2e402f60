"""The fused mining step and its batch-sharded multi-GPU form.

``refine_select`` is one C-ABI call (uem_mine_refine_select_f32) that enqueues, with no host sync,
    [batch max superpixel id] -> [region max of soft] -> [Pearson 1/dist at feature res] ->
    [fused full-res refine (+ per-class max statistics)] -> [pseudo-label selection]
which is tools/train_ssl_uem.py:209-214 / vis_corrected_pseudo_labels.py:185-189 of the reference.

``ShardedMiner`` shards the target batch (or tile list) over the ranks of a torch.distributed job
(one process per GPU).  Per-pixel / per-image / per-region work is rank-local; only
  * the batch-global max superpixel id (alignment.py:241)           -> all_reduce(MAX), 8 B
  * prototype partial sums (c,k) + counts (c) (alignment.py:347-353) -> all_reduce(SUM), <= 57 KiB
  * class histogram (c+1) (balance.py:49-52)                         -> all_reduce(SUM), 64 B
cross the NVLink fabric.
"""
import ctypes

import torch

from . import _lib as L
from . import config, ops


class StatsCache:
    """Class statistics of a refined map handed from label_refine to pseudo_selection (side channel on the tensor).
    ``valid_for`` checks everything the selection kernel relies on, so a stale or foreign cache is never trusted."""
    __slots__ = ("stats", "version", "b", "c", "device")

    def __init__(self, stats, version, b, c, device):
        self.stats, self.version, self.b, self.c, self.device = stats, version, b, c, device

    def valid_for(self, mask):
        return (mask.dim() == 4 and mask._version == self.version and mask.shape[0] == self.b and mask.shape[1] == self.c
                and mask.device == self.device and mask.dtype == torch.float32 and self.stats.device == mask.device
                and tuple(self.stats.shape) == (self.b, self.c + 2))


def _raise_on_status(ws, what):
    """The chain records ids it could not place in the status word at ws[0..3] (include/uem_b200.h): bit 2 = a superpixel
    id outside [0, R) (such pixels are routed to the all-ones sentinel row, i.e. NOT refined by the superpixel view).
    The reference's scatter / gather raises on those (alignment.py:245-250), so under config.strict_asserts this does too
    (one 4-byte device->host read); the word is cleared so a persistent workspace does not carry stale bits.  With
    strict_asserts off nothing is read back: out-of-range ids silently take the sentinel row."""
    word = ws[:4].view(torch.int32)
    bits = int(word.item())
    if bits:
        word.zero_()
        if bits & 2:
            raise RuntimeError("%s: superpixel id outside [0, num_regions) (index out of bounds in the reference's "
                               "scatter/gather, alignment.py:245-250); raise Aligner.num_regions" % what)
        raise RuntimeError("%s: label outside {ignore} U [0, class_num)" % what)


def mine_workspace(soft, num_regions, h, w, k):
    """Zero-initialised scratch of the fused chain for one batch shape (self-cleaning: allocate once, reuse)."""
    b, c, H, W = soft.shape
    need = L.bind(soft).uem_mine_ws_bytes(b, c, H, W, max(h, 1), max(w, 1), max(k, 1), max(int(num_regions), 1))
    return torch.zeros(max(int(need), 16), dtype=torch.uint8, device=soft.device)


def region_phase(soft, sup, temp, num_regions, ws, h, w, k, exchange=None):
    """The region half of the chain on its own (multi-GPU form): region maxima of ``soft`` -> superpixel-view weights in
    ``ws``; returns the rank-LOCAL max superpixel id as a (1,) int64 view INTO ``ws`` (read it -- e.g. pack it for the
    exchange -- before the matching ``refine_select(..., regions_ready=True)``, whose selection kernel zeroes the slot).
    Nothing here needs the batch-global id, so this can run one step ahead of the exchange (alignment.py:241-258).
    exchange: None or (PeerExchange, slot, global_id_out | None): the region-max kernel's last CTA then also sends the id
    part of the step's exchange (``PeerExchange.send(part="id")`` without a launch of its own)."""
    L.require_cuda(soft, sup, ws)
    soft = L.f32c(soft.detach())
    sup = L.i64c(sup.detach())
    b, c, H, W = soft.shape
    R = int(num_regions)
    lib = L.bind(soft)
    dims = (b, c, H, W, max(h, 1), max(w, 1), max(k, 1), max(R, 1))
    assert ws.numel() >= lib.uem_mine_ws_bytes(*dims), "workspace too small: mining.mine_workspace(...)"
    if exchange is not None:
        peer, slot, global_id_out = exchange
        assert (peer.c, peer.k) == (c, dims[6]), "the exchange region was sized for another (class_num, feat_channels)"
        L.require_cuda(global_id_out)
        L.check(lib.uem_mine_region_phase_xchg_f32(L.ptr(sup), R, L.ptr(soft), b, c, H, W, dims[4], dims[5], dims[6], ops.f32(temp),
                                                   L.ptr(ws), peer._arr, peer.rank, peer.world, peer.depth, int(slot),
                                                   L.ptr(global_id_out), L.stream_of(soft)))
    else:
        L.check(lib.uem_mine_region_phase_f32(L.ptr(sup), R, L.ptr(soft), b, c, H, W, dims[4], dims[5], dims[6], ops.f32(temp),
                                              L.ptr(ws), L.stream_of(soft)))
    off = lib.uem_mine_ws_maxid_offset(*dims)
    return ws[off:off + 8].view(torch.int64)


def proto_phase(feat, prototypes, soft_shape, num_regions, ws, eps=1e-7):
    """The prototype half of the chain on its own: 1/Pearson distance of ``feat`` (b,k,h,w) to the prototype bank into the
    similarity slot of ``ws`` (alignment.py:215-217).  It depends on the prototype bank only, so a pipelined loop runs it as
    soon as the previous step's EMA is done, next to that step's refine / selection kernels; the matching
    ``refine_select(..., simi_ready=True)`` then skips it."""
    L.require_cuda(feat, prototypes, ws)
    feat = L.f32c(feat.detach())
    prototypes = L.f32c(prototypes.detach())
    b, c, H, W = soft_shape
    _, k, h, w = feat.shape
    assert prototypes.shape == (c, k) and feat.shape[0] == b
    lib = L.bind(feat)
    R = max(int(num_regions), 1)
    assert ws.numel() >= lib.uem_mine_ws_bytes(b, c, H, W, h, w, k, R), "workspace too small: mining.mine_workspace(...)"
    L.check(lib.uem_mine_proto_phase_f32(L.ptr(feat), k, L.ptr(prototypes), b, c, H, W, h, w, R, ops.f32(eps), L.ptr(ws),
                                         L.stream_of(feat)))


def refine_select(views, soft, temp, feat=None, prototypes=None, pred1=None, pred2=None, sup=None, num_regions=None,
                  ignored_id=None, eps=1e-7, select=None, ws=None, uvem=None, want_entropy=False, regions_ready=False,
                  simi_ready=False):
    """Returns (refined (b,c,H,W) fp32, hard (b,H,W) int64 or None); with ``uvem``/``want_entropy`` the tuple
    grows to (refined, hard, entropy (b*H*W,) or None, uvem_weight (b*H*W,) or None), all from the same pass.

    select: None or (cutoff_top, cutoff_low, ignore_label).
    uvem: None or (m, threshold, gamma) of UVEMLoss (balance.py:345-423); needs ``select``.
    num_regions: capacity R of the region table (ids must lie in [0,R)); None -> read sup.max() back
        (one 8-byte device->host copy, like torch_scatter's ``int(index.max())+1``).
    ignored_id: optional (1,) int64 device tensor holding the batch-global max id (e.g. after an
        all_reduce(MAX) across ranks); None -> computed from ``sup`` inside the call.
    regions_ready: ``region_phase`` already ran on ``ws`` for this batch (needs ``num_regions``, ``ignored_id``, ``ws``).
    simi_ready: ``proto_phase`` already ran on ``ws`` for this batch (``feat`` is still passed: it carries the shapes)."""
    L.require_cuda(soft, feat, prototypes, pred1, pred2, sup, ignored_id)
    soft = L.f32c(soft.detach())
    b, c, H, W = soft.shape
    lib = L.bind(soft)
    k = h = w = 0
    if views & ops.VIEW_PROTO:
        feat = L.f32c(feat.detach())
        prototypes = L.f32c(prototypes.detach())
        _, k, h, w = feat.shape
        assert prototypes.shape == (c, k), "prototypes must be (class_num, feat_channels)"
    if views & ops.VIEW_PRED:
        pred1 = L.f32c(pred1.detach())
        pred2 = None if pred2 is None else L.f32c(pred2.detach())
        if h:
            assert pred1.shape[-2:] == (h, w), "feature map and logits must share their resolution"
        h, w = pred1.shape[-2:]
    R = int(num_regions) if num_regions is not None else 0  # the workspace layout depends on (b, c, R): keep R per workspace
    if views & ops.VIEW_SUP:
        sup = L.i64c(sup.detach())
        assert sup.numel() == b * H * W
        if num_regions is None:
            mm = ops.i64_minmax(sup)
            R = int(mm[1].item()) + 1
            if ignored_id is None:
                ignored_id = mm[1:]
        else:
            R = int(num_regions)
    need = lib.uem_mine_ws_bytes(b, c, H, W, max(h, 1), max(w, 1), max(k, 1), max(R, 1))
    if ws is None or ws.numel() < need:
        # the chain keeps its scratch self-cleaning: it must be zero the first time (include/uem_b200.h)
        ws = torch.zeros(max(int(need), 16), dtype=torch.uint8, device=soft.device)
    refined = torch.empty_like(soft)
    hard = None
    top = low = 0.0
    ign = -1
    if select is not None:
        top, low, ign = select
        hard = torch.empty((b, H, W), dtype=torch.int64, device=soft.device)
    ent = wgt = uv = None
    if want_entropy or uvem is not None:
        assert select is not None, "entropy / UVEM weight are produced by the selection pass"
        ent = torch.empty(b * H * W, dtype=torch.float32, device=soft.device)
        if uvem is not None:
            wgt = torch.empty(b * H * W, dtype=torch.float32, device=soft.device)
            uv = (ctypes.c_float * 5)(*ops._uvem_coefs(*uvem))
    if regions_ready:
        assert num_regions is not None and ignored_id is not None and ws is not None and (views & ops.VIEW_SUP)
        views = int(views) | ops.VIEW_REGIONS_READY
    if simi_ready:
        assert ws is not None and (views & ops.VIEW_PROTO)
        views = int(views) | ops.VIEW_SIMI_READY
    L.check(lib.uem_mine_refine_select_f32(
        int(views), L.ptr(feat), k, L.ptr(prototypes), L.ptr(pred1), L.ptr(pred2), h, w, L.ptr(sup), R, L.ptr(ignored_id),
        L.ptr(soft), b, c, H, W, ops.f32(temp), ops.f32(eps), ops.f32(top), ops.f32(low), int(ign), L.ptr(refined),
        L.ptr(hard), uv, L.ptr(ent), L.ptr(wgt), L.ptr(ws), L.stream_of(soft)))
    if config.strict_asserts and (int(views) & ops.VIEW_SUP):
        _raise_on_status(ws, "label_refine")
    # [class maxima | -min | bad] of `refined`, reused by pseudo_selection() so it needs no second max pass
    if select is None:  # the drop-in pairing label_refine -> pseudo_selection: hand the class statistics over
        # (every path of the C call leaves the table of THIS call in the workspace: uem_refine.cu, end of the chain)
        off = lib.uem_mine_ws_stats_offset(b, c, H, W, max(h, 1), max(w, 1), max(k, 1), max(R, 1))
        stats = ws[off:off + b * (c + 2) * 4].view(torch.int32).view(b, c + 2).clone()
        refined._uem_stats = StatsCache(stats, refined._version, b, c, refined.device)
    if want_entropy or uvem is not None:
        return refined, hard, ent, wgt
    return refined, hard


def mine_step(aligner, label_t_sup, feat_t, preds_t, label_t_soft, mode="all", temp=2.0, cutoff_top=0.8, cutoff_low=0.6,
              ignore_label=-1, ws=None):
    """label_refine + pseudo_selection of one target batch as a single fused call (no host sync when
    aligner.num_regions is set).  Returns (label_t_soft_refined, label_t_hard)."""
    views = ops.MODE_VIEWS[mode]
    pred1, pred2 = (preds_t if isinstance(preds_t, (list, tuple)) else (preds_t, None))
    return refine_select(views, label_t_soft, temp, feat=feat_t, prototypes=aligner.prototypes, pred1=pred1, pred2=pred2,
                         sup=label_t_sup, num_regions=aligner.num_regions, eps=aligner.eps,
                         select=(cutoff_top, cutoff_low, ignore_label), ws=ws)


# ----------------------------------------------------------------------------------------------------
# multi-GPU: batch sharding, tiny all-reduces only
# ----------------------------------------------------------------------------------------------------
def shard_range(n_items, rank, world_size):
    """Contiguous slice [lo,hi) of n_items owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_stats(sums, counts, hist=None):
    """[c*k sums | c counts | c+1 hist] as one fp64 vector (exact for the integer parts) for a single all_reduce."""
    parts = [sums.reshape(-1).double(), counts.reshape(-1).double()]
    if hist is not None:
        parts.append(hist.reshape(-1).double())
    return torch.cat(parts)


def unpack_stats(buf, c, k, with_hist):
    sums = buf[:c * k].float().reshape(c, k)
    counts = buf[c * k:c * k + c].round().long()
    hist = buf[c * k + c:c * k + 2 * c + 1].round().long() if with_hist else None
    return sums, counts, hist


def pack_local(sums, counts, max_id):
    """[c*k sums | c counts | max superpixel id] as one fp64 vector: the rank-local statistics of one step.
    Every part is exact in fp64 (fp32 sums, integer counts and ids < 2^53)."""
    if sums.is_cuda:
        return ops.pack_local(sums, counts, max_id)   # one launch
    return torch.cat([sums.reshape(-1).double(), counts.reshape(-1).double(), max_id.reshape(-1)[:1].double()])


def fold_gathered(gathered, c, k):
    """(world, c*k+c+1) all-gathered rank vectors -> (sums (c,k) fp32, counts (c,) int64, global max id (1,) int64).
    The sum runs over ranks in rank order on every rank, so all ranks get bit-identical prototypes."""
    if gathered.is_cuda:
        return ops.fold_gathered(gathered.contiguous(), c, k)   # one launch
    tot = gathered[0].clone()
    for r in range(1, gathered.shape[0]):
        tot[:c * k + c] += gathered[r, :c * k + c]
    sums = tot[:c * k].float().reshape(c, k)
    counts = tot[c * k:c * k + c].round().long()
    max_id = gathered[:, c * k + c].max().round().long().reshape(1)
    return sums, counts, max_id


class ShardedMiner:
    """Runs the mining step on this rank's shard of a global batch and keeps the replicated state
    (prototype bank, class-frequency EMA) identical on every rank.

    Two ways to drive it:
      * eager: ``mine`` / ``update_prototype`` (all_reduce per exchange);
      * three-phase, for CUDA-graph replay with the collective kept OUTSIDE the captured regions:
        ``local_stats`` (capturable) -> ``exchange`` (one all_gather of <= 100 KB, eager) -> ``apply`` (capturable)."""

    def __init__(self, aligner, group=None, exchange="nccl", depth=3):
        """exchange: 'nccl' (host-issued collectives only), 'peer' (device-side peer stores, uemda_b200/exchange.py; raises
        if the regions cannot be mapped) or 'auto' (peer when it can be set up, else nccl).  depth: exchange slots."""
        import torch.distributed as dist
        self.dist = dist
        self.aligner = aligner
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.peer = None
        if exchange in ("peer", "auto") and self.world > 1:
            from .exchange import ExchangeError, PeerExchange
            try:
                self.peer = PeerExchange.create(aligner.class_num, aligner.prototypes.shape[1], depth=depth, group=group,
                                                device=aligner.prototypes.device)
            except ExchangeError:
                if exchange == "peer":
                    raise

    # ---- device-side form (one or two CUDA graphs per step, no host-issued collective).  Per step and slot, in stream order:
    #   region_phase_send (or region_phase + send_stats(part="id"))  ->  exchange_apply (or send_stats(part="sums") + apply_peer)
    def region_phase_send(self, soft, sup, temp, num_regions, ws, h, w, slot, global_id_out):
        """Capturable.  ``region_phase`` of this rank's shard whose region-max kernel also sends the rank-local max id to
        every rank (slot ``slot``) and leaves the batch-global ignored id (alignment.py:241) in ``global_id_out`` -- the id
        part of the step's exchange without a launch of its own.  Returns the rank-local id like ``region_phase``."""
        return region_phase(soft, sup, temp, num_regions, ws, h, w, self.peer.k, exchange=(self.peer, slot, global_id_out))

    def send_stats(self, partials, local_max_id, slot, hist=None, global_id_out=None, part="all"):
        """Capturable.  partials: ops.proto_accumulate(feat_s, down, ..., fold=False); local_max_id: what region_phase
        returned; hist: optional (c+1,) class histogram of this rank's labels.  Stores this rank's vector into slot
        ``slot`` of every rank; with ``global_id_out`` the same launch waits for the other ranks' ids and leaves the
        batch-global ignored id there (then ``receive_id`` is not needed).  part="id" right after the region pass and
        part="sums" once the source statistics exist split the send so that neither sits at the tail of a step."""
        self.peer.send(partials, local_max_id, slot, hist=hist, global_id_out=global_id_out, part=part)

    def receive_id(self, slot, out=None):
        """Capturable.  Blocks the stream until every rank's vector has arrived -> batch-global ignored id (alignment.py:241)."""
        return self.peer.wait_max_id(slot, out=out)

    def apply_peer(self, slot, in_place=True, want_hist=False):
        """Capturable, after receive_id on the same stream and after the last reader of the old prototype bank has been
        enqueued: rank-ordered fold + EMA of the prototype bank (alignment.py:347-353); returns the global class
        histogram when asked (balance.py:45-52)."""
        al = self.aligner
        new, _, _, hist = self.peer.fold_finalize(slot, al.prototypes, eps=al.eps, decay=al.decay,
                                                  out=al.prototypes if in_place else None, want_hist=want_hist)
        al.prototypes = new
        return hist

    def exchange_apply(self, partials, slot, in_place=True, hist=None, want_hist=False):
        """Capturable: ``send_stats(part="sums")`` + ``apply_peer`` in ONE launch (alignment.py:347-353 over the global batch);
        the step's max id has been sent with ``send_stats(None, local_max_id, slot, part="id")``."""
        al = self.aligner
        new, _, _, h = self.peer.exchange_fold(partials, slot, al.prototypes, eps=al.eps, decay=al.decay,
                                               out=al.prototypes if in_place else None, hist=hist, want_hist=want_hist)
        al.prototypes = new
        return h

    # ---- class histogram over the global batch (balance.py:45-52)
    def class_hist(self, label_local, class_num=None, ignore_label=None):
        """(c+1,) int64: per-class pixel counts and the valid count over the GLOBAL batch: local histogram ->
        all_reduce(SUM).  ClassBalance._local_freq is batch-global in the un-sharded reference."""
        al = self.aligner
        hist = ops.class_hist(label_local, al.class_num if class_num is None else class_num,
                              al.ignore_label if ignore_label is None else ignore_label)
        if self.world > 1:
            self.dist.all_reduce(hist, op=self.dist.ReduceOp.SUM, group=self.group)
        return hist

    def shard_class_balancer(self, balancer):
        """Makes a ClassBalance (uemda_b200.gast.balance) all-reduce its label histograms over this miner's group, so a
        sharded UVEMLoss(class_balancer=...) sees the frequencies of the whole batch like the un-sharded reference."""
        balancer.shard(self.dist if self.world > 1 else None, self.group)
        return balancer

    # ---- three-phase form
    def local_stats(self, sup_local, feat_s_local, label_s_local, out=None, local_max_id=None):
        """Rank-local [prototype sums | counts | max superpixel id] of this step -> fp64 vector (written into ``out``
        when given).  Also returns the down-scaled source labels.  local_max_id: the (1,) int64 tensor ``region_phase``
        returned for this batch (then no separate pass over the ids is needed)."""
        al = self.aligner
        down = al.downscale_gt(label_s_local)
        sums, counts = ops.proto_accumulate(feat_s_local, down, al.class_num, al.ignore_label)
        mx = local_max_id if local_max_id is not None else ops.i64_minmax(sup_local)[1:]
        if out is not None and sums.is_cuda:
            packed = ops.pack_local(sums, counts, mx, out=out)
        else:
            packed = pack_local(sums, counts, mx)
            if out is not None:
                out.copy_(packed)
                packed = out
        return packed, down

    def exchange(self, packed, out=None):
        """all_gather of the packed rank vectors -> (world, n) fp64."""
        if out is None:
            out = torch.empty((self.world, packed.numel()), dtype=packed.dtype, device=packed.device)
        if self.world > 1:
            self.dist.all_gather_into_tensor(out.reshape(-1), packed, group=self.group)
        else:
            out[0].copy_(packed)
        return out

    def fold(self, gathered):
        """Gathered statistics folded in rank order -> (sums, counts, batch-global ignored id (1,) int64)."""
        al = self.aligner
        return fold_gathered(gathered, al.class_num, al.prototypes.shape[1])

    def apply(self, gathered, in_place=False):
        """fold + EMA update of the prototype bank (alignment.py:347-353); returns the ignored id (alignment.py:241).
        in_place: write the new bank over the old tensor (only after every reader of the old bank has been enqueued)."""
        al = self.aligner
        sums, counts, max_id = self.fold(gathered)
        _, new = ops.proto_finalize(sums, counts, al.prototypes, eps=al.eps, decay=al.decay, want_local=False,
                                    out=al.prototypes if in_place else None)
        al.prototypes = new
        return max_id

    def global_ignored_id(self, sup_local):
        """all_reduce(MAX) of the local max superpixel id -> (1,) int64 device tensor (alignment.py:241)."""
        mx = ops.i64_minmax(sup_local)[1:].clone()
        if self.world > 1:
            self.dist.all_reduce(mx, op=self.dist.ReduceOp.MAX, group=self.group)
        return mx

    def mine(self, sup_local, feat_local, preds_local, soft_local, mode="all", temp=2.0, cutoff_top=0.8, cutoff_low=0.6,
             ignore_label=-1, num_regions=None):
        views = ops.MODE_VIEWS[mode]
        ignored = None
        if views & ops.VIEW_SUP:
            ignored = self.global_ignored_id(sup_local)
            if num_regions is None:
                num_regions = int(ignored.item()) + 1
        pred1, pred2 = (preds_local if isinstance(preds_local, (list, tuple)) else (preds_local, None))
        return refine_select(views, soft_local, temp, feat=feat_local, prototypes=self.aligner.prototypes, pred1=pred1,
                             pred2=pred2, sup=sup_local, num_regions=num_regions, ignored_id=ignored, eps=self.aligner.eps,
                             select=(cutoff_top, cutoff_low, ignore_label))

    def update_prototype(self, feat_local, label_local):
        """Aligner.update_prototype over the global batch: local masked sums -> all_reduce(SUM) -> EMA."""
        al = self.aligner
        down = al.downscale_gt(label_local)
        sums, counts = ops.proto_accumulate(feat_local, down, al.class_num, al.ignore_label)
        if self.world > 1:
            buf = pack_stats(sums, counts)
            self.dist.all_reduce(buf, op=self.dist.ReduceOp.SUM, group=self.group)
            sums, counts, _ = unpack_stats(buf, al.class_num, sums.shape[1], False)
        _, new = ops.proto_finalize(sums, counts, al.prototypes, eps=al.eps, decay=al.decay, want_local=False)
        al.prototypes = new
        return down
