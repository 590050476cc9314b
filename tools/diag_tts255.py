"""r02 diagnostic: clip 255 of the full-size configs[3] test differs from the oracle by 1.7e-4 on the fp16-split tcgen05
conv path (1.4e-5 on the fp32 FFMA path; alone == in batch, so it is arithmetic).  Where does it grow?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nisqa_b200 import engine as E, synth
from oracle import nisqa_oracle as O
import test_gpu_parity as T
args, sd = O.load_checkpoint(os.path.join(ROOT, "weights", "nisqa_tts.tar"))
eng = E.Engine(E.config_from_args(args), 0); eng.load_state_dict(sd)
eng.set_option("keep_td_out", 1); eng.set_option("conv12", 0)
clips = T._sliced_clips([10.0] * 256, 16000, seed=6)
stages = [("mel_db", E.STAGE_MEL_DB), ("pool1", E.STAGE_POOL1), ("pool2", E.STAGE_POOL2), ("conv3", E.STAGE_CONV3),
          ("pool3", E.STAGE_POOL3), ("conv5", E.STAGE_CONV5), ("cnn_feat", E.STAGE_CNN_FEAT), ("td_out", E.STAGE_TD_OUT)]
for i in (255, 254):
    taps = {}
    ref, ns, st = O.predict_pcm(args, sd, clips[i].astype(np.float32) / 32768.0, 16000, taps)
    d = {}
    for mode, opts in (("tc", dict(conv_tc=1, conv_split=1)), ("ffma", dict(conv_tc=0, conv_split=0))):
        for k, v in opts.items():
            eng.set_option(k, v)
        sc = eng.predict_pcm([clips[i]], [16000])[0]
        d[mode] = (float(sc[0, 0]), {n: eng.stage_dump(s) for n, s in stages})
    print("clip", i, "score tc %.7f ffma %.7f oracle %.7f" % (d["tc"][0], d["ffma"][0], float(ref[0])))
    for n, _ in stages:
        r = np.asarray(taps[n].numpy() if hasattr(taps[n], "numpy") else taps[n], dtype=np.float32).reshape(-1)
        a, b = d["tc"][1][n], d["ffma"][1][n]
        k = int(np.argmax(np.abs(a - b)))
        print("  %-9s max|ref| %9.3f  tc-ffma %.3e (at value %.4g)  tc-oracle %.3e  ffma-oracle %.3e  min nonzero |ref| %.3e" % (
            n, float(np.abs(r).max()), float(np.abs(a - b).max()), float(b[k]), float(np.abs(a - r).max()), float(np.abs(b - r).max()),
            float(np.abs(r[r != 0]).min()) if np.any(r != 0) else 0.0))
    t = d["tc"][1]["td_out"].reshape(-1, 256); f = d["ffma"][1]["td_out"].reshape(-1, 256)
    dd = np.abs(t - f).max(axis=1)
    print("  td_out row-wise max diff: first rows", dd[:3], "middle", dd[490:493], "last", dd[-3:], "argmax row", int(dd.argmax()))
