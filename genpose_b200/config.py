"""Flag surface of the reference's single argparse namespace (configs/config.py:8-84): same flag names,
types and defaults, so scripts/eval_single.sh and the runners parse unchanged.  Table-driven; flags that
only training/data code reads are still accepted (and ignored by genpose_b200)."""
import argparse

_STORE_TRUE = object()

# (flag, type-or-_STORE_TRUE, default, extra kwargs)
_FLAGS = [
    # dataset
    ("synset_names", str, ["bottle", "bowl", "camera", "can", "laptop", "mug"], {"nargs": "+"}),
    ("selected_classes", str, None, {"nargs": "+"}),
    ("data_path", str, None, {}),
    ("o2c_pose", _STORE_TRUE, True, {}),
    ("batch_size", int, 192, {}), ("max_batch_size", int, 192, {}), ("mini_bs", int, 192, {}),
    ("pose_mode", str, "rot_matrix", {}), ("seed", int, 0, {}),
    ("percentage_data_for_train", float, 1.0, {}), ("percentage_data_for_val", float, 1.0, {}),
    ("percentage_data_for_test", float, 1.0, {}),
    ("train_source", str, "CAMERA+Real", {}), ("val_source", str, "CAMERA", {}), ("test_source", str, "Real", {}),
    ("device", str, "cuda", {}), ("num_points", int, 1024, {}), ("per_obj", str, "", {}), ("num_workers", int, 32, {}),
    # model
    ("posenet_mode", str, "score", {}), ("hidden_dim", int, 128, {}),
    ("sampler_mode", str, None, {"nargs": "+"}), ("sampling_steps", int, None, {}),
    ("sde_mode", str, "ve", {}), ("sigma", float, 25, {}), ("likelihood_weighting", _STORE_TRUE, False, {}),
    ("regression_head", str, "Rx_Ry_and_T", {}), ("pointnet2_params", str, "light", {}), ("pts_encoder", str, "pointnet2", {}),
    ("energy_mode", str, "IP", {}), ("s_theta_mode", str, "score", {}), ("norm_energy", str, "identical", {}),
    # training (accepted, unused)
    ("agent_type", str, "score", {}), ("pretrained_score_model_path", str, None, {}),
    ("pretrained_energy_model_path", str, None, {}), ("distillation", _STORE_TRUE, False, {}),
    ("n_epochs", int, 1000, {}), ("log_dir", str, "debug", {}), ("optimizer", str, "Adam", {}), ("eval_freq", int, 100, {}),
    ("repeat_num", int, 20, {}), ("grad_clip", float, 1.0, {}), ("ema_rate", float, 0.999, {}), ("lr", float, 1e-3, {}),
    ("warmup", int, 100, {}), ("lr_decay", float, 0.98, {}), ("use_pretrain", _STORE_TRUE, False, {}),
    ("parallel", _STORE_TRUE, False, {}), ("num_gpu", int, 4, {}), ("is_train", _STORE_TRUE, False, {}),
    # testing
    ("eval", _STORE_TRUE, False, {}), ("pred", _STORE_TRUE, False, {}), ("model_name", str, None, {}),
    ("eval_repeat_num", int, 50, {}), ("save_video", _STORE_TRUE, False, {}), ("max_eval_num", int, 10000000, {}),
    ("results_path", str, "", {}), ("T0", float, 1.0, {}),
    # nocs_mrcnn testing
    ("img_size", int, 256, {}), ("result_dir", str, "", {}), ("model_dir_list", str, None, {"nargs": "+"}),
    ("energy_model_dir", str, "", {}), ("score_model_dir", str, "", {}), ("ranker", str, "energy_ranker", {}),
    ("pooling_mode", str, "nearest", {}),
    # genpose_b200 extension (not in the reference): 'philox' in-kernel noise or 'torch' generator-compatible noise
    ("noise_mode", str, "philox", {}),
    # genpose_b200 extension: 'auto' | 'bf16x3' (tcgen05 tensor cores, error-compensated) | 'fp32' (FFMA parity kernel)
    ("precision", str, "auto", {}),
]


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser()
    for name, typ, default, extra in _FLAGS:
        if typ is _STORE_TRUE:
            parser.add_argument(f"--{name}", default=default, action="store_true")
        elif default is None:
            parser.add_argument(f"--{name}", type=typ, **extra)
        else:
            parser.add_argument(f"--{name}", type=typ, default=default, **extra)
    return parser


def get_config(argv=None):
    cfg = build_parser().parse_args(argv)
    # augmentation parameter dicts the reference attaches (configs/config.py:89-110); data code only
    cfg.DYNAMIC_ZOOM_IN_PARAMS = {"DZI_PAD_SCALE": 1.5, "DZI_TYPE": "uniform", "DZI_SCALE_RATIO": 0.25, "DZI_SHIFT_RATIO": 0.25}
    cfg.PTS_AUG_PARAMS = {"aug_pc_pro": 0.2, "aug_pc_r": 0.2, "aug_rt_pro": 0.3, "aug_bb_pro": 0.3, "aug_bc_pro": 0.3}
    cfg.DEFORM_2D_PARAMS = {"roi_mask_r": 3, "roi_mask_pro": 0.5}
    return cfg
