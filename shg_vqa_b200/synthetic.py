"""Seeded synthetic AGQA/STAR-shaped inputs.

Shapes and value ranges follow what the reference's data layer hands the model:
question ids ``[CLS] w.. [SEP] 0..`` padded to 40 (entry.py:50-63), ``boxes`` = ones
[B,393] float64 (agqa_data.py:216-217), relation/action segment ids = frame index
repeated per slot (entry.py:83), ragged per-frame label vectors
(entry.py:85-86), one-hot answers.  CPU generator ⇒ identical on every box.
"""
import torch

N_FRAMES = 16
N_REL = 8
N_ACT = 3
SEQ_LEN = 40
N_VIS = 393


def make_inputs(batch, seed=1234, n_rel_cls=457, n_act_cls=158, n_ans=171, full_frames=False,
                with_feat=True, visn_tokens=False):
    g = torch.Generator().manual_seed(seed)
    inp = {}
    if with_feat:
        inp["feat"] = torch.randn(batch, 2048, 16, 7, 7, generator=g).relu_()
    if visn_tokens:
        inp["visn_tokens"] = torch.randn(batch, N_VIS, 768, generator=g)
    inp["boxes"] = torch.ones(batch, N_VIS, dtype=torch.float64)
    ids = torch.zeros(batch, SEQ_LEN, dtype=torch.int64)
    mask = torch.zeros(batch, SEQ_LEN, dtype=torch.int64)
    for b in range(batch):
        n = int(torch.randint(6, 29, (1,), generator=g))
        ids[b, 0] = 101
        ids[b, 1:1 + n] = torch.randint(1000, 30522, (n,), generator=g)
        ids[b, 1 + n] = 102
        mask[b, :n + 2] = 1
    inp["input_ids"], inp["input_mask"] = ids, mask
    inp["segment_ids"] = torch.zeros_like(ids)
    inp["rel_segment_ids"] = torch.arange(N_FRAMES).repeat_interleave(N_REL).expand(batch, -1).contiguous()
    inp["act_segment_ids"] = torch.arange(N_FRAMES).repeat_interleave(N_ACT).expand(batch, -1).contiguous()

    def ragged(per_frame, n_cls):
        tokens = torch.zeros(batch, N_FRAMES, per_frame, dtype=torch.int64)
        lengths = torch.zeros(batch, N_FRAMES, dtype=torch.int64)
        for b in range(batch):
            for f in range(N_FRAMES):
                n = per_frame if full_frames else int(torch.randint(0, per_frame + 1, (1,), generator=g))
                lengths[b, f] = n
                tokens[b, f, :n] = torch.randint(1, n_cls, (n,), generator=g)
        return tokens, lengths
    inp["rel_tokens"], inp["rel_lengths"] = ragged(N_REL, n_rel_cls)
    inp["act_tokens"], inp["act_lengths"] = ragged(N_ACT, n_act_cls)
    inp["rel_labels"] = frame_label_list(inp["rel_tokens"], inp["rel_lengths"])
    inp["act_labels"] = frame_label_list(inp["act_tokens"], inp["act_lengths"])
    # hg_mask = (tokens > 0), actions first then relations per frame (agqa_data.py:248-252)
    inp["hg_mask"] = torch.cat([(inp["act_tokens"] > 0), (inp["rel_tokens"] > 0)], dim=2).to(torch.int64)
    ans = torch.randint(0, n_ans, (batch,), generator=g)
    inp["target"] = torch.zeros(batch, n_ans).scatter_(1, ans[:, None], 1.0)
    return inp


def frame_label_list(tokens, lengths):
    """[B,16,q_f] padded tokens + [B,16] lengths → flat list of B*16 ragged int64
    vectors, the order HungarianMatcher flattens them in (matcher.py:65)."""
    b, t, _ = tokens.shape
    return [tokens[i, f, : int(lengths[i, f])].clone() for i in range(b) for f in range(t)]
