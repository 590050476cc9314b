"""Front-end kernel time vs. frame pairs per CTA (engine option fe_ppc) on the bench workload."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from nisqa_b200 import engine as E, synth
from oracle import nisqa_oracle as O
args, sd = O.load_checkpoint(os.path.join(ROOT, "weights", "nisqa.tar"))
eng = E.Engine(E.config_from_args(args), 0); eng.load_state_dict(sd)
base = synth.synth_speech_pcm16(7, 12.0, 48000)
clips = [np.roll(base, 977 * i)[:480000].copy() for i in range(64)]
srs = [48000] * 64
for _ in range(20):
    eng.predict_pcm(clips, srs)
eng.set_profiling(True)
for ppc in [int(v) for v in (sys.argv[1:] or "1 2 4 6 8 12 16 25 32 63 125 250".split())]:
    eng.set_option("fe_ppc", ppc)
    t = []
    for _ in range(12):
        eng.predict_pcm(clips, srs)
        t.append(eng.group_ms("frontend"))
    t = sorted(t)
    print("ppc %3d  frontend median %.4f ms  min %.4f" % (ppc, t[len(t) // 2], t[0]), flush=True)
