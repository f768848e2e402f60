"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A function-by-function CPU restatement (torch CPU ops, fp32) of the reference's
pseudo-label mining path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the
product package ``uemda_b200`` never does and has no CPU fallback.

Pinning status: the reference ships no tests, golden vectors or known-answer fixtures for
this path.  The oracle is therefore pinned against OUTPUTS OF
THE REFERENCE ITSELF: ``oracle/gen_golden.py`` executes the unmodified reference functions
from /root/reference (under ``oracle/ref_shim.py``) on seeded inputs and commits the
results as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` requires this file to
reproduce them bit-for-bit (integer outputs) / to 0 ulp-level tolerance (float outputs, same
ATen kernels).  The one seam that is a restatement on both sides is ``torch_scatter.scatter``
(third-party, unpinned version, source absent): parity at that seam is "unpinned".

Every function cites the reference lines it follows.  The arithmetic order (which operand is
rounded when) is kept identical to the reference; the structure is not.
"""
import math

import torch
import torch.nn.functional as F

EPS = 1e-7  # Aligner.eps, uemda/gast/alignment.py:44


# --------------------------------------------------------------------------------------
# third-party seam: torch_scatter.scatter (call sites alignment.py:187 'sum', :245 'max')
# --------------------------------------------------------------------------------------
def region_reduce(src, index, reduce, dim_size=None):
    """src (b,N,c); index (b,N,1) int64 broadcast over c; out (b,R,c), R=index.max()+1.
    Untouched slots are 0 (torch_scatter fills 'max' with 0 where nothing was scattered)."""
    b, n, c = src.shape
    idx = index.reshape(b, n, 1).expand(b, n, c)
    R = int(index.max()) + 1 if dim_size is None else dim_size
    out = torch.zeros(b, R, c, dtype=src.dtype)
    if reduce == "sum":
        out.scatter_add_(1, idx, src)
    elif reduce == "max":
        out.scatter_reduce_(1, idx, src, "amax", include_self=False)
    elif reduce == "mean":
        out.scatter_add_(1, idx, src)
        cnt = torch.zeros(b, R, c, dtype=src.dtype).scatter_add_(1, idx, torch.ones_like(src))
        out = out / cnt.clamp(min=1)
    else:
        raise ValueError(reduce)
    return out


# --------------------------------------------------------------------------------------
# a1-a4: softmax confidence / entropy / argmax from logits
# --------------------------------------------------------------------------------------
def upsample_bilinear(x, size):
    """alignment.py:219-220,228-229,233; train_align_uem.py:158-159 -- always align_corners=True."""
    return F.interpolate(x, size, mode="bilinear", align_corners=True)


def soft_from_logits(x1, x2, size):
    """tools/train_align_uem.py:158-160: mean of the two heads' softmax after upsampling."""
    u1 = upsample_bilinear(x1, size)
    if x2 is None:
        return u1.softmax(dim=1)
    u2 = upsample_bilinear(x2, size)
    return (u1.softmax(dim=1) + u2.softmax(dim=1)) * 0.5


def entropy(soft):
    """uemda/gast/balance.py:368,372 (also :331, pseudo_generation.py:185): sum_c -p*log(p);
    p == 0 gives NaN exactly as in the reference.  soft (b,c,H,W) -> (b*H*W,)."""
    c = soft.shape[1]
    flat = soft.permute(0, 2, 3, 1).reshape(-1, c)
    return torch.sum(-flat * torch.log(flat), dim=1)


def confidence_argmax(soft):
    """pseudo_generation.py:47,:148: max prob and first-index argmax over classes."""
    conf, arg = torch.max(soft, dim=1)
    return conf, arg


def uvem_weight(u, m, t, gamma):
    """UVEMLoss.get_weight, balance.py:396-423.  Coefficients are Python doubles applied to
    fp32 tensors; clamp before pow; NaN falls through every comparison."""
    left = torch.ones_like(u)
    if m > 0:
        x = torch.where((u <= m) & (u >= 0), u, left)
        x = (-1 / (m ** 2)) * (x - m) ** 2 + 1
        left = torch.clamp(x, min=0.0, max=1.0) ** (1.0 / gamma)
    right = torch.zeros_like(u)
    if m < t:
        x = torch.where((u > m) & (u <= t), u, torch.zeros_like(u))
        x = (-1 / ((t - m) ** 2)) * (x - m) ** 2 + 1
        right = torch.clamp(x, min=0.0, max=1.0) ** (1.0 / gamma)
    wgt = torch.where(u <= m, left, right)
    return torch.where(u >= t, torch.zeros_like(u), wgt)


# --------------------------------------------------------------------------------------
# a5: class-wise thresholded pseudo labels
# --------------------------------------------------------------------------------------
def class_thresholds(mask, cutoff_top, cutoff_low):
    """pseudo_generation.py:76-81: thr[b,c] = max(fp32(max_px p) * cutoff_top, cutoff_low)."""
    b, c = mask.shape[:2]
    peak = mask.reshape(b, c, -1).max(dim=-1, keepdim=True)[0]
    peak = peak * cutoff_top
    floor = torch.tensor([cutoff_low], dtype=mask.dtype)
    return torch.maximum(peak, floor)  # (b,c,1)


def pseudo_select(mask, cutoff_top=0.8, cutoff_low=0.6, ignore_label=-1):
    """pseudo_generation.py:59-93.  Kept iff exactly one class is strictly above its threshold."""
    assert mask.max() <= 1 and mask.min() >= 0
    b, c, h, w = mask.shape
    flat = mask.reshape(b, c, -1)
    thr = class_thresholds(mask, cutoff_top, cutoff_low)
    wins = (flat > thr)
    n_win = wins.sum(dim=1)
    lab = wins.to(mask.dtype).argmax(dim=1)
    lab[n_win != 1] = ignore_label
    return lab.reshape(b, h, w)


def pseudo_select_v1(mask, cutoff_top=0.8, cutoff_low=0.6, ignore_label=-1):
    """pseudo_generation.py:24-56: argmax label, dropped iff its prob is below its class threshold."""
    assert mask.max() <= 1 and mask.min() >= 0
    b, c, h, w = mask.shape
    flat = mask.reshape(b, c, -1)
    thr = class_thresholds(mask, cutoff_top, cutoff_low).permute(0, 2, 1)  # (b,1,c)
    conf, lab = torch.max(flat, dim=1)
    px_thr = torch.sum(thr * F.one_hot(lab, num_classes=c), dim=-1)
    lab[conf < px_thr] = ignore_label
    return lab.reshape(b, h, w)


# --------------------------------------------------------------------------------------
# a8: DownscaleLabel
# --------------------------------------------------------------------------------------
def downscale_label(label, scale_factor=16, n_classes=7, ignore_label=-1, min_ratio=0.75):
    """alignment.py:494-509: block class histogram (c+1 bins) -> majority (first index) ->
    ignore if majority is the ignore bin or its share < min_ratio.  (b,H,W)|(b,1,H,W) -> (b,1,h,w)."""
    lab = label.clone()
    if lab.dim() == 4:
        lab = lab.squeeze(1)
    lab[lab == ignore_label] = n_classes
    hot = F.one_hot(lab, num_classes=n_classes + 1).permute(0, 3, 1, 2).float()
    share = F.avg_pool2d(hot, kernel_size=scale_factor)
    top, out = torch.max(share, dim=1, keepdim=True)
    out[out == n_classes] = ignore_label
    out[top < min_ratio] = ignore_label
    return out


def index_onehot(label, class_num, ignore_label=-1):
    """alignment.py:468-481 (and balance.py:60-67): (b,1,h,w)|(b,h,w) -> (n,c) int64, ignore -> zero row."""
    lab = label.clone()
    if lab.dim() < 4:
        lab = lab.unsqueeze(1)
    lab = lab.permute(0, 2, 3, 1).reshape(-1)
    lab[lab == ignore_label] = class_num
    return F.one_hot(lab, num_classes=class_num + 1)[:, :-1]


# --------------------------------------------------------------------------------------
# a9: Pearson distance
# --------------------------------------------------------------------------------------
def pearson_dist(f1, f2, eps=EPS):
    """alignment.py:424-451.  (n,k),(m,k) -> (n,m) in [0,1]."""
    k = f1.shape[-1]
    d1 = (f1 - f1.mean(dim=-1, keepdim=True)).unsqueeze(1)
    d2 = (f2 - f2.mean(dim=-1, keepdim=True)).unsqueeze(0)
    if d1.shape[0] * d2.shape[1] * k > (1 << 28):  # chunk the (n,m,k) broadcast; same per-row sums
        cov = torch.cat([(d1[i:i + 4096] * d2).sum(dim=-1) for i in range(0, d1.shape[0], 4096)])
    else:
        cov = (d1 * d2).sum(dim=-1)
    cov = cov / (k - 1 + eps)
    den = f1.std(dim=-1).unsqueeze(1) * f2.std(dim=-1).unsqueeze(0)
    return (-1.0 * cov / (den + eps) + 1.0) * 0.5


# --------------------------------------------------------------------------------------
# a6: label_refine
# --------------------------------------------------------------------------------------
def _softmax_t(x, temp):
    """alignment.py:311-314: division first, then softmax over dim 1."""
    return torch.softmax(x / temp, dim=1)


def _peak_norm(x):
    """alignment.py:222,235,253: divide by (max over classes + 1e-7)."""
    return x / (torch.max(x, dim=1, keepdim=True)[0] + 1e-7)


def proto_view_weight(feat, prototypes, size):
    """alignment.py:213-222: 1/pearson -> (b,c,h,w) -> bilinear up -> softmax(T=1) -> /max."""
    b, k, h, w = feat.shape
    flat = feat.permute(0, 2, 3, 1).reshape(-1, k)
    simi = 1.0 / pearson_dist(flat, prototypes)
    simi = simi.view(b, h, w, -1).permute(0, 3, 1, 2)
    simi = upsample_bilinear(simi, size)
    return _peak_norm(_softmax_t(simi, 1))


def pred_view_weight(preds, size, temp):
    """alignment.py:225-235: one head or exactly two heads averaged."""
    if isinstance(preds, (list, tuple)):
        assert len(preds) == 2
        p = (_softmax_t(upsample_bilinear(preds[0], size), temp)
             + _softmax_t(upsample_bilinear(preds[1], size), temp)) * 0.5
    else:
        p = _softmax_t(upsample_bilinear(preds, size), temp)
    return _peak_norm(p)


def sup_view_weight(sup, soft, temp):
    """alignment.py:240-253: region-wise max of soft (scatter 'max'), gathered back per pixel,
    softmax(/temp), /max.  Also returns the batch-global 'ignored' mask (:241-243)."""
    b, c, H, W = soft.shape
    ids = sup.reshape(b, -1, 1)
    ignored = (ids == ids.max()).reshape(b, 1, H, W)
    rows = soft.permute(0, 2, 3, 1).reshape(b, -1, c)
    table = region_reduce(rows, ids, "max")
    px = torch.gather(table, 1, ids.expand(b, H * W, c)).reshape(b, H, W, c).permute(0, 3, 1, 2)
    return _peak_norm(_softmax_t(px, temp)), ignored


def label_refine(sup, feat, preds, soft, prototypes, mode="all", temp=2.0, refine=True):
    """Aligner.label_refine, alignment.py:194-293 (modes 'all','s','p','l')."""
    assert mode in ("all", "s", "p", "l")
    if not refine:
        return soft
    size = soft.shape[-2:]
    wgt = 0
    if mode in ("all", "p"):
        wgt = wgt + proto_view_weight(feat, prototypes, size)
    if mode in ("all", "l"):
        wgt = wgt + pred_view_weight(preds, size, temp)
    if mode in ("all", "s"):
        sw, ignored = sup_view_weight(sup, soft, temp)
        ignored = ignored.expand_as(sw)
        if mode == "all":
            wgt = torch.where(ignored, wgt, wgt * sw)
        else:
            wgt = torch.where(ignored, torch.ones_like(sw), sw)
    out = wgt * soft
    return out / (out.sum(dim=1, keepdim=True) + EPS)  # _logits_norm, alignment.py:316-326


def proto_weight_4pixel(feat, label_hard, prototypes, class_num, ignore_label=-1):
    """Aligner.get_prototype_weight_4pixel, alignment.py:295-309 -> (b*H*W,)."""
    b = feat.shape[0]
    H, W = label_hard.shape[-2:]
    wmap = proto_view_weight(feat, prototypes, (H, W))
    hot = index_onehot(label_hard, class_num, ignore_label).reshape(b, H, W, -1).permute(0, 3, 1, 2)
    return torch.sum(wmap * hot, dim=1).reshape(-1)


# --------------------------------------------------------------------------------------
# a7: superpixel_expand
# --------------------------------------------------------------------------------------
def superpixel_expand(hard, sup, class_num, ignore_label=-1):
    """alignment.py:175-192: per-region class counts (scatter 'sum' of int64 one-hot), majority
    class = first-index argmax, empty regions -> -1, gathered back to pixels."""
    b, H, W = hard.shape
    hot = index_onehot(hard, class_num, ignore_label).reshape(b, -1, class_num)
    ids = sup.reshape(b, -1, 1)
    counts = region_reduce(hot, ids, "sum")
    top, winner = torch.max(counts, dim=-1, keepdim=True)
    winner[top == 0] = -1
    return torch.gather(winner, 1, ids).reshape(b, H, W)


# --------------------------------------------------------------------------------------
# a10-a12: prototypes
# --------------------------------------------------------------------------------------
def ema(history, curr, decay):
    """alignment.py:463-466."""
    return (1.0 - decay) * curr + decay * history


def local_prototypes(feat, label_down, prototypes, class_num, ignore_label=-1, eps=EPS):
    """alignment.py:340-350: masked per-class feature mean; classes without samples keep the old
    prototype.  Returns (local (c,k), counts (c,) int64)."""
    b, k, h, w = feat.shape
    rows = feat.permute(0, 2, 3, 1).reshape(-1, k)
    hot = index_onehot(label_down, class_num, ignore_label)  # (n,c) int64
    cnt = hot.sum(0)
    sums = torch.zeros(class_num, k)
    for i in range(0, rows.shape[0], 2048):  # chunked (n,c,k) broadcast
        sums += (rows[i:i + 2048].unsqueeze(1) * hot[i:i + 2048].unsqueeze(-1)).sum(0)
    n_inst = cnt.reshape(class_num, 1).expand(class_num, k)
    local = sums / (n_inst + eps)
    return torch.where(n_inst < 1, prototypes, local), cnt


def class_feature_sums(feat, label_down, class_num, ignore_label=-1):
    """alignment.py:109-119 (update_avg): raw masked sums (c,k) and counts (c,1) float."""
    b, k, h, w = feat.shape
    rows = feat.permute(0, 2, 3, 1).reshape(-1, k)
    hot = index_onehot(label_down, class_num, ignore_label)
    sums = torch.zeros(class_num, k)
    for i in range(0, rows.shape[0], 2048):
        sums += (rows[i:i + 2048].unsqueeze(1) * hot[i:i + 2048].unsqueeze(-1)).sum(0)
    return sums, hot.sum(0).reshape(class_num, 1).float()


def soft_prototypes(feat, soft):
    """update_prototype_bytarget, alignment.py:98-104: mean over pixels of feat * down(soft)."""
    b, k, h, w = feat.shape
    c = soft.shape[1]
    rows = feat.permute(0, 2, 3, 1).reshape(-1, 1, k)
    down = F.interpolate(soft, size=(h, w), mode="bilinear", align_corners=True)
    down = down.permute(0, 2, 3, 1).reshape(-1, c, 1)
    acc = torch.zeros(c, k)
    for i in range(0, rows.shape[0], 2048):
        acc += (rows[i:i + 2048] * down[i:i + 2048]).sum(0)
    return acc / rows.shape[0]


# --------------------------------------------------------------------------------------
# a14: ClassBalance
# --------------------------------------------------------------------------------------
def flat_onehot(label, class_num, ignore_label=-1):
    """ClassBalance._one_hot, balance.py:60-67: any label shape -> (numel, c) int64, ignore -> zero row."""
    lab = label.clone().reshape(-1)
    lab[lab == ignore_label] = class_num
    return F.one_hot(lab, num_classes=class_num + 1)[:, :-1]


def class_histogram(label, class_num, ignore_label=-1):
    """balance.py:45-52,60-67: per-class pixel counts and number of non-ignored pixels."""
    hot = flat_onehot(label, class_num, ignore_label)
    valid = torch.sum((label != ignore_label).float())
    return hot.sum(dim=0), valid


def class_balance_step(freq, label, class_num, ignore_label=-1, decay=0.99, temperature=0.5, eps=EPS):
    """ClassBalance.get_class_weight_4pixel, balance.py:27-43: EMA update of freq FIRST, then the
    weight table softmax((1-freq)/T)/max, then the per-pixel lookup (ignored -> 0)."""
    hist, valid = class_histogram(label, class_num, ignore_label)
    local = hist.float() / (valid + eps)
    freq = (1.0 - decay) * local + decay * freq
    prob = torch.softmax((1.0 - freq) / temperature, dim=0)
    table = prob / (torch.max(prob, dim=0, keepdim=True)[0] + eps)
    hot = flat_onehot(label, class_num, ignore_label)
    weight = (hot * table.unsqueeze(0)).sum(dim=1)
    return freq, table, weight


# --------------------------------------------------------------------------------------
# a15: generic histograms
# --------------------------------------------------------------------------------------
def hist_f32(x, bins, lo, hi):
    """torch.histc as used at balance.py:193 (30 bins over [0,1])."""
    return torch.histc(x, bins=bins, min=lo, max=hi)


# --------------------------------------------------------------------------------------
# whole mining step (the unit bench.py times): tools/train_ssl_uem.py:209-216 + balance.py:372-396
# --------------------------------------------------------------------------------------
def mining_step(inp, prototypes, class_num, mode="all", temp=2.0, cutoff_top=0.8, cutoff_low=0.6,
                decay=0.996, uvem=(0.2, 0.7, 4.0), ignore_label=-1, scale_factor=16):
    refined = label_refine(inp["sup"], inp["feat"], [inp["pred1"], inp["pred2"]], inp["soft"],
                           prototypes, mode=mode, temp=temp)
    hard = pseudo_select(refined, cutoff_top, cutoff_low, ignore_label)
    out = dict(refined=refined, hard=hard)
    if "feat_s" in inp:
        down = downscale_label(inp["label_s"], scale_factor, class_num, ignore_label, 0.75)
        local, cnt = local_prototypes(inp["feat_s"], down, prototypes, class_num, ignore_label)
        out.update(label_s_down=down, local=local, counts=cnt, prototypes=ema(prototypes, local, decay))
    u = entropy(refined)
    out.update(entropy=u, uvem_weight=uvem_weight(u, *uvem))
    return out


# --------------------------------------------------------------------------------------
# next row 8f-2: PrototypeContrastiveLoss (uemda/loss.py:18-47), autograd supplies the backward
# --------------------------------------------------------------------------------------
def pcl_loss(protos, feat, labels, temperature=8.0, ignore_label=-1):
    """feat (b,k,h,w) or (N,k) with requires_grad; returns the scalar loss (call .backward() for d/dfeat)."""
    if feat.dim() != 2:
        k = feat.size(1)
        feat = feat.permute(0, 2, 3, 1).reshape(-1, k)      # loss.py:30-31
    labels = labels.reshape(-1)                             # :32-33
    mask = labels != ignore_label                            # :36-38
    labels = labels[mask]
    feat = feat[mask]
    feat = F.normalize(feat, p=2, dim=1)                     # :40-41
    protos = F.normalize(protos, p=2, dim=1)
    logits = feat.mm(protos.permute(1, 0).contiguous()) / temperature   # :43-44
    return F.cross_entropy(logits, labels)                   # :46


# --------------------------------------------------------------------------------------
# f3: UVEM / UPS target loss over up-sampled heads
# --------------------------------------------------------------------------------------
def uvem_loss_calc(heads, label, label_soft, m=0.1, threshold=0.7, gamma=8.0, use_weight=True, balance_freq=None,
                   class_num=None, ignore_label=-1, multi=True):
    """loss_calc_uvem (balance.py:437-457) over UVEMLoss.forward (:356-394) or, with use_weight=False,
    UPSLoss.forward (:321-342).  heads: list of (b,c,h,w) logits (autograd flows into them); balance_freq: None or the
    ClassBalance frequency vector, which moves once PER HEAD exactly as in the reference (the loss is called per head).
    Returns (loss, final balance_freq)."""
    heads = list(heads) if multi else [heads]
    c = label_soft.shape[1]
    class_num = class_num or c
    t_ = label.long().reshape(-1)
    u = entropy(label_soft).detach()
    total = 0
    for p in heads:
        if p.shape[-2:] != label.shape[-2:]:
            p = upsample_bilinear(p, label.shape[-2:])
        p_ = p.permute(0, 2, 3, 1).reshape(-1, c)
        ce = F.cross_entropy(p_, t_, reduction="none", ignore_index=ignore_label)
        ce = torch.where(u > threshold, torch.zeros_like(ce), ce)
        weight = uvem_weight(u, m, threshold, gamma) if use_weight else 1.0
        if balance_freq is not None:
            balance_freq, _, cw = class_balance_step(balance_freq, t_, class_num, ignore_label)
            weight = weight * cw
        valid = torch.sum((u <= threshold) & (t_ != ignore_label))
        total = total + (weight * ce).sum() / (valid + 1e-7)
    return (total / len(heads) if multi else total), balance_freq


# --------------------------------------------------------------------------------------
# f4: IAST class-wise percentile thresholds + sliding-window accumulation (uemda/utils/tools.py)
# --------------------------------------------------------------------------------------
def ias_thresh(conf_dict, n_class, alpha, w=None, gamma=1.0):
    """tools.py:323-333: per class, np.percentile (linear interpolation) of its confidence list at
    100 * (1 - alpha * w_c ** gamma); classes whose list is None keep 1.0.  Returns float32 (n_class,)."""
    import numpy as np
    if w is None:
        w = np.ones(n_class)
    cls_thresh = np.ones(n_class, dtype=np.float32)
    for c in range(n_class):
        if conf_dict[c] is not None:
            cls_thresh[c] = np.percentile(np.array(conf_dict[c]), 100 * (1 - alpha * w[c] ** gamma))
    return cls_thresh


def iast_batch_step(probs, cls_thresh, alpha, beta, gamma):
    """One batch of generate_pseudo, tools.py:347-371, file IO and visualisation left out.
    probs (b,c,H,W) float32 torch tensor (the model output the reference calls `logits`), cls_thresh float64 (c,) ndarray
    (np.ones(c) * 0.9 before the first batch).  Returns (new cls_thresh float64 (c,), labels uint8 (b,H,W) = class + 1,
    0 where the winning confidence is below its class threshold)."""
    import numpy as np
    n_class = probs.shape[1]
    max_items = probs.max(dim=1)                               # :349  (lowest index on ties)
    label_pred = max_items[1].numpy()
    logits_pred = max_items[0].numpy()
    conf = {c: [cls_thresh[c]] for c in range(n_class)}        # :353  the previous threshold is one sample of the list
    for c in range(n_class):
        conf[c].extend(logits_pred[label_pred == c].astype(np.float16))   # :355  float16!
    tmp = ias_thresh(conf, n_class, alpha, w=cls_thresh, gamma=gamma)     # :357
    cls_thresh = beta * cls_thresh + (1 - beta) * tmp          # :360  float64 * float + float32 * float -> float64
    cls_thresh[cls_thresh >= 1] = 0.999                        # :361
    np_logits = probs.numpy()
    out = []
    for i in range(np_logits.shape[0]):
        logit = np_logits[i].transpose(1, 2, 0)                # :366
        label = np.argmax(logit, axis=2)
        amax = np.amax(logit, axis=2)
        ignore = amax < cls_thresh[label]                      # :369-370 (float32 < float64)
        label = label + 1
        label[ignore] = 0
        out.append(label.astype(np.uint8))
    return cls_thresh, np.stack(out)


def slide_windows(image_hw, tile_size=(512, 512)):
    """Window origins of pre_slide, tools.py:62-80: overlap 1/2, windows clamped to the image.  Returns [(y1, x1, y2, x2)]."""
    from math import ceil
    H, W = image_hw
    stride = ceil(tile_size[0] * (1 - 1 / 2))
    rows = int(ceil((H - tile_size[0]) / stride) + 1)
    cols = int(ceil((W - tile_size[1]) / stride) + 1)
    wins = []
    for r in range(rows):
        for c in range(cols):
            x1, y1 = int(c * stride), int(r * stride)
            x2, y2 = min(x1 + tile_size[1], W), min(y1 + tile_size[0], H)
            x1, y1 = max(int(x2 - tile_size[1]), 0), max(int(y2 - tile_size[0]), 0)
            wins.append((y1, x1, y2, x2))
    return wins


def slide_accumulate(tiles, wins, image_hw):
    """pre_slide, tools.py:69-97 with the model call factored out: tiles[i] (b,c,th,tw) is the (padded) prediction of window
    i; the full map is the per-pixel mean of the windows covering it (sum in window order, then / count)."""
    b, c = tiles[0].shape[:2]
    H, W = image_hw
    full = torch.zeros(b, c, H, W)
    cnt = torch.zeros(b, 1, H, W)
    for t, (y1, x1, y2, x2) in zip(tiles, wins):
        full[:, :, y1:y2, x1:x2] += t[:, :, 0:y2 - y1, 0:x2 - x1]
        cnt[:, :, y1:y2, x1:x2] += 1
    full /= cnt
    return full


def tta_mean(preds):
    """tta_predict, tools.py:132-152 after de-augmentation: mean over the stacked views (cat on dim 0, mean dim 0)."""
    return torch.mean(torch.cat(preds, 0), dim=0, keepdim=True)
