"""GPU-box A/B of the tcgen05 conv variants (one process per variant so that a faulting variant cannot
take the others down):

    python tools/tc_ab.py [--lib path.so] [--split 0|1] [--timing] [--skip-check] [--tag name]

Prints (1) max |variant - fp32 FFMA path| per CNN stage and the worst |score - oracle| on a few clips,
(2) per-layer kernel times (CUDA events on the engine stream) and the device-resident throughput of
64 x 10 s clips, (3) with --timing (library built with -DNISQA_TC_TIMING) the per-phase clock64 stamps."""
import argparse, ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=None)
ap.add_argument("--split", type=int, default=1, help="1: fp16 plane pipeline (conv_split.cu), 0: legacy fp32-activation tcgen05 kernels")
ap.add_argument("--timing", action="store_true")
ap.add_argument("--tag", default="")
ap.add_argument("--skip-check", action="store_true")
a = ap.parse_args()
from nisqa_b200 import engine as E, synth
if a.lib:
    E.load_library(os.path.join(ROOT, a.lib))
    E._lib = E.load_library(os.path.join(ROOT, a.lib))
import torch
from oracle import nisqa_oracle as O
tag = a.tag or ("split=%d lib=%s" % (a.split, a.lib or "default"))


def opts(eng):
    eng.set_option("conv_split", a.split)
    eng.set_option("conv12", 0)          # the per-stage comparison dumps pool1


if not a.skip_check:
    for ckpt, spec in (("nisqa.tar", [(1, 3.0, 48000), (2, 1.37, 48000), (3, 2.0, 16000), (5, 0.1875, 8000)]),
                       ("nisqa_tts.tar", [(10, 2.0, 16000), (11, 1.3, 48000)])):
        args, sd = O.load_checkpoint(os.path.join(ROOT, "weights", ckpt))
        eng = E.Engine(E.config_from_args(args), 0); eng.load_state_dict(sd)
        pcm = [synth.synth_speech_pcm16(s, sec, sr) for (s, sec, sr) in spec]
        srs = [c[2] for c in spec]
        res = {}
        for tc in (0, 1):
            eng.set_option("conv_tc", tc)
            if tc:
                opts(eng)
            sc, nseg, st = eng.predict_pcm(pcm, srs)
            res[tc] = [sc.copy()] + [eng.stage_dump(s) for s in (E.STAGE_POOL1, E.STAGE_POOL2, E.STAGE_CONV3, E.STAGE_POOL3, E.STAGE_CONV5, E.STAGE_CNN_FEAT)]
        d = ["%s %.2e" % (nm, np.abs(res[0][i] - res[1][i]).max()) for i, nm in enumerate(["scores", "pool1", "pool2", "conv3", "pool3", "conv5", "feat"])]
        worst = 0.0
        for i, (p, sr) in enumerate(zip(pcm, srs)):
            ref, ns, st = O.predict_pcm(args, sd, p.astype(np.float32) / 32768.0, sr)
            worst = max(worst, float(np.abs(res[1][0][i] - ref).max()))
        print("[%s] %s  max|tc-ffma|: %s | worst |score-oracle| %.2e" % (tag, ckpt, "  ".join(d), worst), flush=True)
        eng.close()

args, sd = O.load_checkpoint(os.path.join(ROOT, "weights", "nisqa.tar"))
eng = E.Engine(E.config_from_args(args), 0); eng.load_state_dict(sd)
opts(eng)
BS, SR = 64, 48000
base = [synth.synth_speech_pcm16(i, 10.0, SR) for i in range(8)]
clips = [np.roll(base[i % 8], 977 * i) for i in range(BS)]
stride = (len(clips[0]) + 15) // 16 * 16
buf = np.zeros(BS * stride, np.int16)
for i, c in enumerate(clips):
    buf[i * stride:i * stride + len(c)] = c
offs = np.arange(BS, dtype=np.int64) * stride
ns = np.full(BS, len(clips[0]), np.int64); srs = np.full(BS, SR, np.int32)
pcm = [torch.from_numpy(np.roll(buf, 16 * k)).cuda() for k in range(4)]
out = torch.empty((BS, 5), device="cuda")


def step(i, sync=False):
    eng.predict_pcm_device(pcm[i % 4].data_ptr(), offs, ns, srs, E.FMT_S16, out.data_ptr(), sync=sync)


def measure(tag):
    t0 = time.time(); i = 0
    while time.time() - t0 < 1.0:
        step(i); i += 1
    torch.cuda.synchronize()
    K = 150
    t0 = time.perf_counter()
    for i in range(K):
        step(i)
    eng.join(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    eng.set_profiling(True)
    names = ["frontend", "conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "lin_ln", "qkv", "sa_layer", "pool"]
    acc = dict((k, 0.0) for k in names)
    for i in range(5):
        step(i, sync=True)
        for k in names:
            acc[k] += max(eng.group_ms(k), 0.0) / 5
    eng.set_profiling(False)
    print("[%s] %.0f clips/s (%.3f ms/step)  kernels ms: %s" % (tag, K * BS / dt, dt / K * 1e3, "  ".join("%s %.3f" % (k, acc[k]) for k in names)), flush=True)


measure(tag)

if a.timing:
    lib = E.load_library()
    reader = lib.nisqa_debug_sp_timing if a.split else lib.nisqa_debug_tc_timing
    reader.argtypes = [C.POINTER(C.c_longlong), C.c_int]
    for layer in (4, 3, 5, 6, 2):
        if a.split:
            eng.set_option("tc_timing_layer", layer)
        else:
            if layer == 6:
                continue
            eng.set_option("conv_tc", 1 << layer)
        for _ in range(3):
            step(0, sync=True)
        n = 8 * 8192
        b = (C.c_longlong * n)()
        assert reader(b, n) == 0
        t = np.array(b[:], dtype=np.int64).reshape(-1, 8)
        nb = {4: 5270, 3: 5270, 5: 1757, 6: 1757, 2: 15808}[layer]
        t = t[300:min(nb, 8192)]
        d = lambda x, y: float(np.median(t[:, y] - t[:, x]))
        print("[%s] conv%d  setup %.0f  fill %.0f  fill_end->first_mma %.0f  mma_issue %.0f  (fill_end->acc_ready %.0f)  epilogue %.0f  teardown %.0f  total %.0f cycles"
              % (tag, layer, d(0, 1), d(1, 2), d(2, 6), d(6, 7), d(2, 3), d(3, 4), d(4, 5), d(0, 5)), flush=True)
eng.close()
