"""GPU-box experiment: per-phase clock64 stamps of the LAST tcgen05 conv launch (conv6 runs last, so
use conv_tc mask to leave only one layer on the TC path).  Build with -DNISQA_TC_TIMING."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nisqa_b200 import engine as E, synth
from oracle import nisqa_oracle as O
args, sd = O.load_checkpoint(os.path.join(ROOT, "weights", "nisqa.tar"))
eng = E.Engine(E.config_from_args(args), 0); eng.load_state_dict(sd)
clips = [np.roll(synth.synth_speech_pcm16(i % 4, 10.0, 48000), 1000 * i) for i in range(64)]
lib = E.load_library()
for layer in (4, 3, 5, 2):
    eng.set_option("conv_tc", 1 << layer)
    for _ in range(3):
        eng.predict_pcm(clips, [48000] * 64)
    n = 8 * 8192
    buf = (C.c_longlong * n)()
    assert lib.nisqa_debug_tc_timing(buf, n) == 0
    t = np.array(buf[:], dtype=np.int64).reshape(-1, 8)
    nb = {4: 5270, 3: 5270, 5: 1757, 2: 15808}[layer]
    t = t[:min(nb, 8192)]
    t = t[300:]            # skip the first wave
    d = lambda a, b: float(np.median(t[:, b] - t[:, a]))
    print("conv%d  setup %.0f  fill %.0f  fill_end->first_mma %.0f  mma_issue %.0f  (fill_end->acc_ready %.0f)  epilogue %.0f  teardown %.0f  total %.0f cycles"
          % (layer, d(0, 1), d(1, 2), d(2, 6), d(6, 7), d(2, 3), d(3, 4), d(4, 5), d(0, 5)))
