"""Architecture variants reachable through user-trained checkpoints for which the reference ships
no weights: the other pooling modules (PoolAtt / PoolAvg / PoolMax / PoolLastStep, reference lib:1117-1225) and the
positional encoding of the self-attention block (lib:1042-1062).

TEST INFRASTRUCTURE ONLY (like the rest of oracle/).  A variant checkpoint = a shipped checkpoint's args with the
variant's options switched + its state_dict with the pooling head replaced by deterministic seeded weights (NumPy
PCG64: the same here, in the golden generator that feeds them to the UNMODIFIED reference modules, and on the GPU box).
"""
import math

import numpy as np
import torch

# name -> (base checkpoint, args overrides)
VARIANTS = {
    "dim_pool_avg": ("nisqa.tar", {"pool": "avg", "pool_att_h": None}),
    "dim_pool_max": ("nisqa.tar", {"pool": "max", "pool_att_h": None}),
    "dim_pool_last_step": ("nisqa.tar", {"pool": "last_step", "pool_att_h": None}),
    "dim_pool_att": ("nisqa.tar", {"pool": "att", "pool_att_h": None}),
    "dim_pos_enc": ("nisqa.tar", {"td_sa_pos_enc": True}),
    "mos_pool_att_pos_enc": ("nisqa_mos_only.tar", {"pool": "att", "pool_att_h": None, "td_sa_pos_enc": True}),
    "tts_pool_avg": ("nisqa_tts.tar", {"pool": "avg"}),
    "tts_pool_max": ("nisqa_tts.tar", {"pool": "max"}),
    "tts_pool_last_step": ("nisqa_tts.tar", {"pool": "last_step"}),
}
# clips every variant is scored on: (seed, seconds, sample rate)
CLIPS = [(71, 2.0, 48000), (72, 0.9, 16000), (73, 3.7, 44100)]


def positional_encoding(d_model=64, max_len=3000):
    """The registered buffer of PositionalEncoding.__init__ (reference lib:1051-1058: the PyTorch tutorial formula,
    same torch ops so that the values are the ones the reference module builds)."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).transpose(0, 1).contiguous()


def variant_checkpoint(name, base_args, base_sd):
    """-> (args, state_dict) of the variant."""
    _, over = VARIANTS[name]
    args = dict(base_args)
    args.update(over)
    sd = {k: v for k, v in base_sd.items()}
    pool_changed = "pool" in over
    if pool_changed:
        sd = {k: v for k, v in sd.items() if not (k.startswith("pool.") or k.startswith("pool_layers."))}
        d = 256 if args["td"] == "lstm" else 64
        heads = ["pool_layers.%d.model." % i for i in range(5)] if args["model"] == "NISQA_DIM" else ["pool.model."]
        rng = np.random.default_rng(sum(map(ord, name)))
        for pf in heads:
            def lin(key):
                sd[pf + key + ".weight"] = torch.from_numpy((rng.standard_normal((1, d)) * 0.15).astype(np.float32))
                sd[pf + key + ".bias"] = torch.from_numpy(rng.uniform(1.0, 4.0, 1).astype(np.float32))
            if args["pool"] == "att":
                lin("linear1")
                lin("linear2")
            else:
                lin("linear")
    if args.get("td_sa_pos_enc"):
        sd["time_dependency.model.pos_encoder.pe"] = positional_encoding()
    return args, sd
