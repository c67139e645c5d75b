#!/bin/bash
# hang / race hunting at small grids for every tile-team size and both arithmetics, then compute-sanitizer memcheck on a short case
OUT=gpurun_out; mkdir -p $OUT
for P in f16x2 bf16x3; do for TEAM in 1 2 4; do
  echo "== stress $P team $TEAM"; timeout 300 python tools/tc_stress.py 24 $P $TEAM 2>&1 | tail -4
done; done | tee $OUT/r2k_tc_stress.txt
echo "== memcheck (PC T=6 + ODE, teams 1/2/4, f16x2)"
cat > /tmp/mc.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from genpose_b200 import ops, synth
from genpose_b200.sde import init_sde
sd = synth.make_state_dict(1, kappa=0.3); eng = ops.Engine(sd)
B, K = 5, 50
pts = torch.from_numpy(synth.make_clouds(B, 1)).cuda(); cen = pts.mean(dim=1).contiguous()
ob = eng.object_bias(eng.encode(pts))
x0 = torch.from_numpy(synth.make_prior_noise(B * K, 1)).cuda()
torch.manual_seed(0); x0o = init_sde("ve")[0]((B * K, 9), T=0.15).cuda().contiguous()
for team in (1, 2, 4):
    for prec in ("f16x2", "bf16x3"):
        eng.sample_pc(ob, cen, x0, K, 6, seed=1, precision=prec, team=team, return_process=True)
        eng.sample_ode(ob, cen, x0o, K, T0=0.15, precision=prec, team=team, return_process=True)
torch.cuda.synchronize(); print("memcheck workload done")
PY
timeout 1200 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/mc.py 2>&1 | tail -12 | tee $OUT/r2k_memcheck.txt
echo "== phase + sweep after the hoist"
timeout 90 python tools/tc_phase_times.py 100 0 f16x2 2>&1 | head -16 | tail -3
timeout 600 python tools/tc_batch_sweep.py 100 64,378 f16x2 0 2>&1 | grep pc
