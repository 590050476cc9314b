"""One nisqa_tts.tar call on short clips (for ncu captures of the StandardCNN / BiLSTM kernels)."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from nisqa_b200 import engine as E, synth
from oracle import nisqa_oracle as O
args, sd = O.load_checkpoint(os.path.join(ROOT, "weights", "nisqa_tts.tar"))
eng = E.Engine(E.config_from_args(args), 0); eng.load_state_dict(sd)
base = synth.synth_speech_pcm16(3, 10.0, 16000)
clips = [np.roll(base, 331 * i) for i in range(148)]
for _ in range(2):
    eng.predict_pcm(clips, [16000] * len(clips))
