"""GPU-box check of the tcgen05 conv path against the oracle and against the FFMA path."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nisqa_b200 import engine as E, synth
from oracle import nisqa_oracle as O

def run(ckpt, clips):
    args, sd = O.load_checkpoint(os.path.join(ROOT, "weights", ckpt))
    eng = E.Engine(E.config_from_args(args), 0)
    eng.load_state_dict(sd)
    pcm = [synth.synth_speech_pcm16(s, sec, sr) for (s, sec, sr) in clips]
    srs = [c[2] for c in clips]
    res = {}
    for tc in (0, 1):
        eng.set_option("conv_tc", tc)
        sc, nseg, st = eng.predict_pcm(pcm, srs)
        res[tc] = (sc.copy(), eng.stage_dump(E.STAGE_POOL2), eng.stage_dump(E.STAGE_CONV3), eng.stage_dump(E.STAGE_POOL3), eng.stage_dump(E.STAGE_CONV5), eng.stage_dump(E.STAGE_CNN_FEAT))
        print(ckpt, "conv_tc=%d" % tc, "scores[0]", sc[0].tolist(), flush=True)
    for i, nm in enumerate(["scores", "pool2", "conv3", "pool3", "conv5", "feat"]):
        a, b = res[0][i], res[1][i]
        print("  %s: max|tc-ffma| = %.3e  (max |ffma| %.3e)" % (nm, np.abs(a - b).max(), np.abs(a).max()))
    worst = 0
    for i, (p, sr) in enumerate(zip(pcm, srs)):
        ref, ns, st = O.predict_pcm(args, sd, p.astype(np.float32) / 32768.0, sr)
        worst = max(worst, np.abs(res[1][0][i] - ref).max())
    print("  worst |tc - oracle| score = %.3e" % worst, flush=True)
    eng.close()

run("nisqa.tar", [(1, 3.0, 48000), (2, 1.37, 48000), (3, 2.0, 16000), (0, 10.0, 48000), (5, 0.1875, 8000)])
run("nisqa_tts.tar", [(10, 2.0, 16000), (11, 1.3, 48000)])
