"""GPU-box probe: nisqa_tts.tar throughput and per-group kernel times at full length (987 segments / clip)."""
import os, sys, numpy as np, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from nisqa_b200 import engine as E, synth
from oracle import nisqa_oracle as O
args, sd = O.load_checkpoint(os.path.join(ROOT, "weights", "nisqa_tts.tar"))
eng = E.Engine(E.config_from_args(args), 0); eng.load_state_dict(sd)
bases = [synth.synth_speech_pcm16(900 + i, 12.0, 16000) for i in range(4)]
for n, kind in ((148, "roll"), (256, "roll"), (256, "slices"), (512, "roll")):
    if kind == "roll":
        clips = [np.roll(bases[0][:160000], 331 * i) for i in range(n)]
    else:
        rng = np.random.default_rng(1)
        clips = [bases[i % 4][int(rng.integers(0, 30000)):][:160000] for i in range(n)]
    eng.set_profiling(False)
    eng.predict_pcm(clips, [16000] * n)
    eng.set_profiling(True)
    eng.predict_pcm(clips, [16000] * n)
    print(n, kind, {g: round(eng.group_ms(g), 3) for g in ("frontend", "cnn", "fc_out", "lstm")}, flush=True)
