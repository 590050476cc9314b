"""Per-kernel times of the bench workload (64 x 10 s 48 kHz clips, nisqa.tar) for one library build - the A/B tool for
variant libraries built by tools/tc_ab_build.sh.

    python tools/ab_kernels.py [--lib nisqa_b200/exp/libnisqa_X.so] [--tag X]
"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser(); ap.add_argument("--lib", default=None); ap.add_argument("--tag", default="default")
a = ap.parse_args()
from nisqa_b200 import engine as E, synth
if a.lib:
    E._lib = E.load_library(os.path.join(ROOT, a.lib))
import torch
ck = torch.load(os.path.join(ROOT, "weights", "nisqa.tar"), map_location="cpu", weights_only=False)
eng = E.Engine(E.config_from_args(ck["args"]), 0); eng.load_state_dict(ck["model_state_dict"])
base = synth.synth_speech_pcm16(7, 12.0, 48000)
clips = [np.roll(base, 977 * i)[:480000].copy() for i in range(64)]
srs = [48000] * 64
for _ in range(25):
    sc = eng.predict_pcm(clips, srs)[0]
eng.set_profiling(True)
names = ["frontend", "conv12", "conv3", "conv4", "conv5", "conv6", "lin_ln", "sa_layer", "pool"]
acc = dict((k, []) for k in names)
for _ in range(15):
    eng.predict_pcm(clips, srs)
    for k in names:
        acc[k].append(eng.group_ms(k))
med = dict((k, float(np.median(v))) for k, v in acc.items())
print("[%s] " % a.tag + "  ".join("%s %.4f" % (k, med[k]) for k in names) + "  sum %.4f  score0 %.6f" % (sum(med.values()), float(sc[0, 0])), flush=True)
