"""GPU-box measurement of the reference-facing surface on real files: nisqaModel.predict() in predict_dir mode over
N synthetic 10 s 48 kHz wav files (page cache warm), plus the ingest alone (native probe + decode into pinned
memory, no GPU work) to show which side bounds it.

    python tools/bench_files.py [N] [workers ...]
"""
import os, sys, time, tempfile, shutil, io, contextlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from nisqa_b200 import synth, wav, NISQA_lib as NL
from nisqa_b200.NISQA_model import nisqaModel
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
WORKERS = [int(x) for x in sys.argv[2:]] or [4, 8, 16, 32]
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
base = [synth.synth_speech_pcm16(i, 10.0, 48000) for i in range(8)]
for i in range(N):
    wav.write_wav_pcm16(os.path.join(d, "c%05d.wav" % i), np.roll(base[i % 8], 977 * i), 48000)
print("files in", d, "host:", os.cpu_count(), "cpus,", open("/sys/fs/cgroup/cpu.max").read().strip() if os.path.exists("/sys/fs/cgroup/cpu.max") else "no cpu.max", flush=True)
for workers in WORKERS:
    args = {"mode": "predict_dir", "pretrained_model": os.path.join(ROOT, "weights", "nisqa.tar"), "data_dir": d,
            "output_dir": None, "tr_bs_val": 64, "tr_num_workers": workers, "ms_channel": None}
    with contextlib.redirect_stdout(io.StringIO()):
        m = nisqaModel(args)
    # ingest alone: the loader of the predict loop, batch after batch, nothing submitted to the GPU
    pool = NL._PinnedPool(4)
    rows = np.arange(N)
    t0 = time.perf_counter()
    for b in range(0, N, 64):
        NL._load_batch(m.ds_val, rows[b:b + 64], pool, (b // 64) % 4, workers)
    t_ing = time.perf_counter() - t0
    for rep in range(2):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            df = m.predict()
        dt = time.perf_counter() - t0
    print("num_workers=%2d: predict_dir %d files bs=64: %6.0f clips/s (%.3f s) | ingest alone (1 feeder): %6.0f clips/s = %.1f GB/s | mos[0]=%.6f"
          % (workers, N, N / dt, dt, N / t_ing, N * 0.96e-3 / t_ing, df["mos_pred"].iloc[0]), flush=True)
    m.model.close()
shutil.rmtree(d)
