"""HuBERT span-mask generator (host side, NumPy global stream) -- the bit-exact object of
reference ``fairseq_code/data_utils.py:20-153`` for the options MelHuBERT uses.

The mask is a tiny (B, T) bool array that must reproduce the reference's legacy
``np.random`` stream draw for draw, so it stays on the host; it is shipped to the device as
one byte per frame and applied by ``mh_mask_rows_f32_to_bf16`` / consumed by
``mh_select_rows``.
"""
import numpy as np


def compute_mask_indices(shape, padding_mask, mask_prob, mask_length, mask_type="static", mask_other=0.0,
                         min_masks=0, no_overlap=False, min_space=0, require_same_masks=True, mask_dropout=0.0,
                         valid_lens=None):
    """``padding_mask``: (B, T) bool array/tensor with True at padded frames, or None.
    ``valid_lens`` may be given instead (avoids the reference's per-row device syncs).
    RNG consumption order is identical to the reference."""
    bsz, all_sz = shape
    rng = np.random
    out = np.zeros((bsz, all_sz), dtype=bool)
    all_num = max(min_masks, int(mask_prob * all_sz / float(mask_length) + rng.rand()))
    if valid_lens is None and padding_mask is not None:
        pm = padding_mask.cpu().numpy() if hasattr(padding_mask, "cpu") else np.asarray(padding_mask)
        valid_lens = all_sz - pm.astype(np.int64).sum(axis=1)
    picks = []
    for b in range(bsz):
        if valid_lens is not None:
            sz = int(valid_lens[b])
            n_span = max(min_masks, int(mask_prob * sz / float(mask_length) + rng.rand()))
        else:
            sz, n_span = all_sz, all_num
        if mask_type == "static":
            lengths = np.full(n_span, mask_length)
        elif mask_type == "uniform":
            lengths = rng.randint(mask_other, mask_length * 2 + 1, size=n_span)
        elif mask_type == "normal":
            lengths = np.array([max(1, int(round(x))) for x in rng.normal(mask_length, mask_other, size=n_span)])
        elif mask_type == "poisson":
            lengths = np.array([int(round(x)) for x in rng.poisson(mask_length, size=n_span)])
        else:
            raise Exception("unknown mask selection " + mask_type)
        if lengths.sum() == 0:
            lengths[0] = min(mask_length, sz - 1)
        if no_overlap:
            idx = _non_overlapping(rng, sz, lengths, min_space)
        else:
            shortest = int(lengths.min())
            if sz - shortest <= n_span:
                shortest = sz - n_span - 1
            starts = rng.choice(sz - shortest, n_span, replace=False)
            idx = np.concatenate([s + np.arange(l) for s, l in zip(starts, lengths)]) if n_span else np.zeros(0, int)
        picks.append(np.unique(idx[idx < sz]))
    fewest = min(len(m) for m in picks)
    for b, idx in enumerate(picks):
        if require_same_masks and len(idx) > fewest:
            idx = rng.choice(idx, fewest, replace=False)
        if mask_dropout > 0:
            holes = np.rint(len(idx) * mask_dropout).astype(int)
            idx = rng.choice(idx, len(idx) - holes, replace=False)
        out[b, idx] = True
    return out


def _non_overlapping(rng, sz, lengths, min_space):
    """The reference's no_overlap branch (data_utils.py:95-122; it crashes on NumPy >= 1.24
    because of ``np.int`` -- same algorithm with ``int``)."""
    chosen = []
    parts = [(0, sz)]
    keep = int(min(lengths))
    for length in sorted((int(x) for x in lengths), reverse=True):
        room = np.array([e - s if e - s >= length + min_space else 0 for s, e in parts], dtype=int)
        if room.sum() == 0:
            break
        c = rng.choice(len(parts), p=room / room.sum())
        s, e = parts.pop(c)
        start = rng.randint(s, e - length)
        chosen.extend(range(start, start + length))
        if start - s - min_space >= keep:
            parts.append((s, start - min_space + 1))
        if e - start - length - min_space > keep:
            parts.append((start + length + min_space, e))
    return np.asarray(chosen, dtype=int)
