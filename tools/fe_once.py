"""A few passes of the bench workload (for ncu captures of single kernels)."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from nisqa_b200 import engine as E, synth
from oracle import nisqa_oracle as O
args, sd = O.load_checkpoint(os.path.join(ROOT, "weights", "nisqa.tar"))
eng = E.Engine(E.config_from_args(args), 0); eng.load_state_dict(sd)
base = synth.synth_speech_pcm16(7, 12.0, 48000)
clips = [np.roll(base, 977 * i)[:480000].copy() for i in range(64)]
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    eng.predict_pcm(clips, [48000] * 64)
