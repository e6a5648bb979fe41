"""Experiment configuration (drop-in for model/video_prediction/config.py:6-134).

Same attribute names and defaults as the reference so existing launch scripts
(`--args key value ...` -> setattr) keep working; the modules of this package read only the
model fields.  Differences: `dtype` defaults to float32 (the kernels compute in fp32; the
reference default is float64, config.py:58) and two optional fields are added --
`align_corners` (False = what the reference code executes under torch >= 1.3, True = its
original torch-1.0.1 semantics, SURVEY.md hard part 1) and `use_cuda_graph`.
"""
import torch


class StoveConfig:
    # experiment
    description = 'unnamed experiment'
    nolog = False
    experiment_dir = './experiments/unsorted'
    checkpoint_path = None
    keep_folder = False
    action_conditioned = None
    random_seed = None
    supairvised = False
    load_encoder = None
    supair_only = False
    supair_grad = True
    debug_test_mode = False

    # data
    traindata = './data/billiards_train.pkl'
    testdata = './data/billiards_test.pkl'
    num_visible, num_rollout, frame_step = 8, 8, 1
    num_episodes = 1000
    num_frames = width = height = num_obj = r = coord_lim = action_space = None
    channels = 1
    debug_add_noise = False

    # optimisation
    batch_size = 256
    cl = 32
    learning_rate, min_learning_rate, debug_anneal_lr = 0.002, 0.0002, 40000.0
    num_epochs = 400
    debug_amsgrad = True
    debug_gradient_clip = True

    # runtime
    device = None
    dtype = torch.float32
    max_threads = 8
    num_workers = 4

    # logging
    debug = True
    n_plot_sequences = 5
    print_every = 100
    plot_every = 1e19
    save_every = 10000
    long_rollout_every = 10000
    visdom = False
    debug_extend_plots = False

    # STOVE
    skip = 2
    transition_lik_std = [0.01, 0.01, 0.01, 0.01]
    debug_fix_supair = True
    debug_match_appearance = False
    debug_no_latents = False

    # action-conditioned
    debug_reward_factor = 15000
    debug_reward_rampup = 20000
    debug_mse = False
    debug_core_appearance = False
    debug_appearance_dim = 3

    # dynamics
    debug_nonlinear = 'relu'
    debug_latent_q_std = 0.04
    debug_xavier = False

    # SPNs / SuPAIR
    debug_bw = True
    patch_height = patch_width = 10
    obj_min_var, obj_max_var = 0.12, 0.35
    bg_min_var, bg_max_var = 0.002, 0.16
    scale_var = pos_var = 0.3
    min_obj_scale, max_obj_scale = 0.1, 0.8
    min_y_scale, max_y_scale = 0.75, 1.25
    obj_pos_bound = 0.9
    obj_spn_num_gauss = obj_spn_num_sums = 10
    overlap_beta = 10.0
    debug_bg_model = False
    debug_obj_spn = False
    debug_simple_bg_var = 0.1
    debug_simple_obj_var = 0.2
    debug_match_objects = '3_only'
    debug_no_reuse = False
    debug_no_velocity = False

    # stove_b200 additions
    align_corners = False
    use_cuda_graph = False

    def __init__(self, **overrides):
        for k, v in overrides.items():
            setattr(self, k, v)
