"""GPU-box measurement of BASELINE.json configs[2] (ragged 2-30 s, bs=512, nisqa.tar) and configs[3]
(nisqa_tts.tar, bs=256, 16 kHz 10 s): clips/s and audio-seconds/s with the PCM resident in HBM."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nisqa_b200 import engine as E, synth
from oracle import nisqa_oracle as O


def run(name, ckpt, durations, sr, steps=8, check=3):
    args, sd = O.load_checkpoint(os.path.join(ROOT, "weights", ckpt))
    eng = E.Engine(E.config_from_args(args), 0); eng.load_state_dict(sd)
    bases = [synth.synth_speech_pcm16(900 + i, 30.0, sr) for i in range(8)]
    rng = np.random.default_rng(5)
    clips = []
    for i, d in enumerate(durations):
        n = int(round(d * sr)); st = int(rng.integers(0, len(bases[0]) - n + 1))
        clips.append(bases[i % 8][st:st + n])
    offs, tot = [], 0
    for c in clips:
        offs.append(tot); tot += (len(c) + 15) // 16 * 16
    buf = np.zeros(tot, np.int16)
    for o, c in zip(offs, clips):
        buf[o:o + len(c)] = c
    d_pcm = torch.from_numpy(buf).cuda()
    out = torch.empty((len(clips), eng.n_out), dtype=torch.float32, device="cuda")
    ns = [len(c) for c in clips]; srs = [sr] * len(clips)
    stream = torch.cuda.ExternalStream(eng.stream())
    t_w = time.time()
    while time.time() - t_w < 1.5:
        eng.predict_pcm_device(d_pcm.data_ptr(), offs, ns, srs, E.FMT_S16, out.data_ptr(), sync=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        nseg, status = eng.predict_pcm_device(d_pcm.data_ptr(), offs, ns, srs, E.FMT_S16, out.data_ptr(), sync=False)
    eng.join()                 # lane 0 waits for the other compute lanes: e1 covers all of them
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    got = out.cpu().numpy()
    worst = 0.0
    for i in np.linspace(0, len(clips) - 1, check).astype(int):
        ref, _, _ = O.predict_pcm(args, sd, clips[i].astype(np.float32) / 32768.0, sr)
        worst = max(worst, float(np.abs(got[i] - ref).max()))
    eng.set_profiling(True)
    eng.predict_pcm_device(d_pcm.data_ptr(), offs, ns, srs, E.FMT_S16, out.data_ptr(), sync=True)
    groups = {g: round(eng.group_ms(g), 3) for g in ("frontend", "cnn", "td", "pool", "lstm", "fc_out", "sa_layer")}
    eng.close()
    res = {"config": name, "clips": len(clips), "segments": int(nseg.sum()), "ms_per_step": ms,
           "clips_per_s": len(clips) / ms * 1e3, "audio_s_per_s": float(sum(durations)) / ms * 1e3,
           "parity_max_abs_vs_oracle": worst, "group_ms": groups, "all_ok": bool((status == 0).all())}
    print(json.dumps(res), flush=True)
    return res


if __name__ == "__main__":
    out = []
    out.append(run("configs[2] ragged 2-30 s bs=512 nisqa.tar 48 kHz", "nisqa.tar", list(synth.ragged_durations(512, 2.0, 30.0, seed=11)), 48000))
    out.append(run("configs[3] nisqa_tts.tar bs=256 16 kHz 10 s", "nisqa_tts.tar", [10.0] * 256, 16000))
    out.append(run("nisqa_mos_only.tar bs=64 48 kHz 10 s", "nisqa_mos_only.tar", [10.0] * 64, 48000))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_configs.json"), "w"), indent=1)
