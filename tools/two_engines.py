"""GPU-box experiment: aggregate device-resident throughput of 1 vs 2 engine instances (= 2 compute
streams with private workspaces) on one GPU, alternating 64-clip steps."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nisqa_b200 import engine as E, synth
from oracle import nisqa_oracle as O
args, sd = O.load_checkpoint(os.path.join(ROOT, "weights", "nisqa.tar"))
BS, SR = 64, 48000
base = [synth.synth_speech_pcm16(i, 10.0, SR) for i in range(8)]
clips = [np.roll(base[i % 8], 977 * i) for i in range(BS)]
stride = (len(clips[0]) + 15) // 16 * 16
buf = np.zeros(BS * stride, np.int16)
for i, c in enumerate(clips): buf[i * stride:i * stride + len(c)] = c
offs = np.arange(BS, dtype=np.int64) * stride
ns = np.full(BS, len(clips[0]), np.int64); srs = np.full(BS, SR, np.int32)
for n_eng in (1, 2, 3):
    engs, pcm, outs = [], [], []
    for _ in range(n_eng):
        e = E.Engine(E.config_from_args(args), 0); e.load_state_dict(sd); engs.append(e)
        pcm.append(torch.from_numpy(buf).cuda()); outs.append(torch.empty((BS, 5), device="cuda"))
    def step(i):
        k = i % n_eng
        engs[k].predict_pcm_device(pcm[k].data_ptr(), offs, ns, srs, E.FMT_S16, outs[k].data_ptr(), sync=False)
    t0 = time.time(); i = 0
    while time.time() - t0 < 1.5: step(i); i += 1
    torch.cuda.synchronize()
    K = 120
    t0 = time.perf_counter()
    for i in range(K): step(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("engines=%d  %.0f clips/s  (%.3f ms per 64-clip step)" % (n_eng, K * BS / dt, dt / K * 1e3), flush=True)
    for e in engs: e.close()
