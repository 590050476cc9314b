"""r02 diagnostic: clip 255 of the full-size configs[3] test differed from the oracle by 1.7e-4 - batch effect or arithmetic?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nisqa_b200 import engine as E, synth
from oracle import nisqa_oracle as O
import test_gpu_parity as T
args, sd = O.load_checkpoint(os.path.join(ROOT, "weights", "nisqa_tts.tar"))
eng = E.Engine(E.config_from_args(args), 0); eng.load_state_dict(sd)
clips = T._sliced_clips([10.0] * 256, 16000, seed=6)
idx = [255, 254, 200, 0]
for lb in (0, 1):
    eng.set_option("lstm_batched", lb)
    full, _, _ = eng.predict_pcm(clips, [16000] * 256)
    print("lstm_batched", lb, "in batch:", [float(full[i, 0]) for i in idx])
    alone = [float(eng.predict_pcm([clips[i]], [16000])[0][0, 0]) for i in idx]
    print("   alone   :", alone)
eng.set_option("lstm_batched", 0)
eng.set_option("conv_tc", 0); eng.set_option("conv_split", 0)
print("ffma convs alone:", [float(eng.predict_pcm([clips[i]], [16000])[0][0, 0]) for i in idx])
print("oracle          :", [float(O.predict_pcm(args, sd, clips[i].astype(np.float32) / 32768.0, 16000)[0][0]) for i in idx])
