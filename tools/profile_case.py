"""One encode of the first C2 bench batch (64 items, ~192 audio-s) on ONE stream, twice (warm-up + the pass ncu captures):
the fixed launch order the ncu recipes under profiles/ rely on. Usage: python tools/profile_case.py [K=8] [mode]
Launches per pass that match front_fused|tcp_gemm|swa_attention|rvq_tc|layernorm: front, SEANet convs 3..13 (11),
8 x (LN, QKV, attention, o_proj, LN, fc1, fc2), downsample, input_proj, RVQ = 71."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tokenize_audio_b200 import synth
from tokenize_audio_b200.encoder import MimiB200Model

K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
bench.select_workload("c2")
model = MimiB200Model(synth.synth_state_dict(0), device="cuda:0")
if len(sys.argv) > 2:
    model.set_mode(int(sys.argv[2]))
model.streams = 1
clips, lengths, batches = bench.make_workload(0)
cl = clips[0]
n = max(len(c) for c in cl)
x = torch.zeros(len(cl), 1, n)
for i, c in enumerate(cl):
    x[i, 0, :len(c)] = torch.from_numpy(c)
x = x.cuda()
lens = [len(c) for c in cl]
for _ in range(2):
    codes = model.encode(x, num_quantizers=K, valid_lengths=lens).audio_codes
    torch.cuda.synchronize()
print("audio_s", sum(lens) / 24000, "computed_s", sum(min(n, -(-l // 1920) * 1920) for l in lens) / 24000, "launches", model.launch_count,
      "checksum", int(codes.sum()))
