#!/bin/bash
# BASELINE configs[4]: the reference's own REAL275 evaluation (scripts/eval_single.sh, unchanged flags) with genpose_b200 dropped in.
# Needs what the reference's README.md ("Download dataset and models") lists and this repository cannot ship:
#   $GENPOSE/results/ckpts/ScoreNet/ckpt_genpose.pth, $GENPOSE/results/ckpts/EnergyNet/ckpt_genpose.pth   (released checkpoints)
#   $GENPOSE/results/evaluation_results/segmentation_results_real_test.pkl                               (pre-segmented REAL275 clouds)
#   $NOCS (--data_path): Real/test, gts/real_test, obj_models/real_test                                  (for compute_mAP)
# The unmodified runner is used: PYTHONPATH puts genpose_b200/dropin (networks.posenet_agent, networks.posenet, networks.reward,
# configs.config, pointnet2_cuda) in front of the checkout; everything else (unpack_data, compute_mAP, plotting) is the reference's.
#
# usage: tools/run_real275.sh <genpose checkout> <NOCS data path> [extra flags, e.g. --precision bf16x3 --noise_mode torch]
#   --precision auto (default)  tcgen05 two-product samplers; bf16x3 = three-product (highest fidelity); fp32 = FFMA parity kernels
#   --noise_mode torch          PC sampler only: draw z1/z2 with torch.randn_like in the reference's order (same CUDA generator stream)
# Expected (GenPose paper, arXiv 2306.10531 Table 1 as recalled in BASELINE.md - verify before quoting; REAL275, K = 50, ODE T0 = 0.55, energy
# ranker, ratio 0.6): 5deg2cm 52.1 / 5deg5cm 60.9 / 10deg2cm 72.4 /
# 10deg5cm 84.0 mAP; candidates are random draws, so runs agree statistically (+-0.5), not bitwise.
set -euo pipefail
GENPOSE=${1:?usage: tools/run_real275.sh <genpose checkout> <NOCS data path> [flags]}
NOCS=${2:?NOCS data path}
shift 2
HERE=$(cd "$(dirname "$0")/.." && pwd)
for f in results/ckpts/ScoreNet/ckpt_genpose.pth results/ckpts/EnergyNet/ckpt_genpose.pth results/evaluation_results/segmentation_results_real_test.pkl; do
  [ -f "$GENPOSE/$f" ] || { echo "missing $GENPOSE/$f (reference README: 'Download dataset and models')"; exit 2; }
done
python -c "import __graft_entry__ as g; g.build()" >/dev/null
cd "$GENPOSE"
export PYTHONPATH="$HERE/genpose_b200/dropin:$HERE${PYTHONPATH:+:$PYTHONPATH}"
# one GPU: exactly scripts/eval_single.sh
python runners/evaluation_single.py \
  --score_model_dir ScoreNet/ckpt_genpose.pth --energy_model_dir EnergyNet/ckpt_genpose.pth --data_path "$NOCS" \
  --sampler_mode ode --max_eval_num 1000000 --percentage_data_for_test 1.0 --batch_size 256 --seed 0 --test_source real_test \
  --result_dir results --eval_repeat_num 50 --pooling_mode average --ranker energy_ranker --T0 0.55 "$@"
# results/evaluation_results/real_test_repeat_50/results/average/energy_ranker/eval_logs.txt holds the mAP table (5deg2cm first).
# 8 GPUs: the runner itself is single-process; shard the category batches with genpose_b200.distributed.run_sharded
# or run one process per category subset (--synset_names) and merge the pickles before compute_mAP.
