#!/bin/bash
# GPU-box memory / race / init check of every kernel on the default path (the reference has no
# sanitizer story; this is ours): compute-sanitizer memcheck over __graft_entry__.smoke() (nisqa.tar, two
# short clips at 48 / 16 kHz: front-end, conv1, the five tcgen05 conv layers with their bulk copies /
# mbarriers / TMEM, self-attention, pooling) and over a short nisqa_tts.tar call (StandardCNN, BiLSTM).
#   gpurun --timeout 300 -- 'bash tools/sanitize.sh r01e'
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
cat > /tmp/san_run.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import __graft_entry__ as g
g.smoke()
from nisqa_b200 import engine as E, synth
from oracle import nisqa_oracle as O
args, sd = O.load_checkpoint(os.path.join("weights", "nisqa_tts.tar"))
eng = E.Engine(E.config_from_args(args, max_chunk_segments=64), 0); eng.load_state_dict(sd)   # two passes: shrinking planes
sc, ns, st = eng.predict_pcm([synth.synth_speech_pcm16(7, 0.6, 16000), synth.synth_speech_pcm16(8, 0.3, 16000)], [16000, 16000])
ref, _, _ = O.predict_pcm(args, sd, synth.synth_speech_pcm16(7, 0.6, 16000).astype(np.float32) / 32768.0, 16000)
assert abs(float(sc[0, 0]) - float(ref[0])) <= 1e-4
print("tts ok", sc.tolist())
PY
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file $OUT/${TAG}_memcheck.log python /tmp/san_run.py > $OUT/${TAG}_memcheck_stdout.log 2>&1
echo "memcheck exit $?"
tail -5 $OUT/${TAG}_memcheck.log; tail -3 $OUT/${TAG}_memcheck_stdout.log
# shared-memory hazards (racecheck) and uninitialised device reads (initcheck) of the same run
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file $OUT/${TAG}_racecheck.log python /tmp/san_run.py > $OUT/${TAG}_racecheck_stdout.log 2>&1
echo "racecheck exit $?"
tail -4 $OUT/${TAG}_racecheck.log
timeout 280 compute-sanitizer --tool initcheck --error-exitcode 7 --log-file $OUT/${TAG}_initcheck.log python /tmp/san_run.py > $OUT/${TAG}_initcheck_stdout.log 2>&1
echo "initcheck exit $?"
tail -4 $OUT/${TAG}_initcheck.log
