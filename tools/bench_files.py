"""GPU-box measurement of the reference-facing surface on real files: nisqaModel.predict() in
predict_dir mode over N synthetic 10 s 48 kHz wav files on local disk (page cache warm)."""
import os, sys, time, tempfile, shutil
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from nisqa_b200 import synth, wav
from nisqa_b200.NISQA_model import nisqaModel
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
d = tempfile.mkdtemp()
base = [synth.synth_speech_pcm16(i, 10.0, 48000) for i in range(8)]
for i in range(N):
    wav.write_wav_pcm16(os.path.join(d, "c%05d.wav" % i), np.roll(base[i % 8], 977 * i), 48000)
for workers in (4, 16, 32):
    args = {"mode": "predict_dir", "pretrained_model": os.path.join(ROOT, "weights", "nisqa.tar"), "data_dir": d,
            "output_dir": None, "tr_bs_val": 64, "tr_num_workers": workers, "ms_channel": None}
    m = nisqaModel(args)
    import io, contextlib
    for rep in range(2):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            df = m.predict()
        dt = time.perf_counter() - t0
    print("predict_dir %d files, bs=64, num_workers=%d: %.0f clips/s (%.3f s)  mos[0]=%.6f" % (N, workers, N / dt, dt, df["mos_pred"].iloc[0]), flush=True)
shutil.rmtree(d)
