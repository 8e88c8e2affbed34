"""The reference's command-line flags (nerf_sh/nerf/utils.py:60-253, octree/nerf/utils.py:60-253) as absl flags with
the same names and defaults, and `update_flags` (utils.py:233-244): a YAML file `<config>.yaml` overrides them, so the
reference's own config files (nerf_sh/config/blender.yaml, tt.yaml) drive this package unchanged.

Flags that select features outside the scope of this path are accepted (so existing command lines keep parsing) and
rejected by `check_scope` when set to an unsupported value."""
import os

import yaml
from absl import flags

FLAGS = flags.FLAGS

# name -> (kind, default, help)
_COMMON = {
    "train_dir": ("string", None, "where to store ckpts and logs"),
    "data_dir": ("string", None, "input data directory."),
    "config": ("string", None, "using config files to set hyperparameters."),
    "dataset": ("string", "blender", "The type of dataset feed to nerf."),
    "image_batching": ("bool", False, "sample rays in a batch from different images."),
    "white_bkgd": ("bool", True, "using white color as default background."),
    "batch_size": ("integer", 1024, "the number of rays in a mini-batch (for training)."),
    "factor": ("integer", 4, "the downsample factor of images, 0 for no downsample."),
    "spherify": ("bool", False, "set for spherical 360 scenes."),
    "render_path": ("bool", False, "render generated path if set true."),
    "llffhold": ("integer", 8, "will take every 1/N images as LLFF test set."),
    "model": ("string", "nerf", "name of model to use."),
    "near": ("float", 2.0, "near clip of volumetric rendering."),
    "far": ("float", 6.0, "far clip of volumentric rendering."),
    "net_depth": ("integer", 8, "depth of the first part of MLP."),
    "net_width": ("integer", 256, "width of the first part of MLP."),
    "net_depth_condition": ("integer", 1, "depth of the second part of MLP."),
    "net_width_condition": ("integer", 128, "width of the second part of MLP."),
    "weight_decay_mult": ("float", 0.0, "The multiplier on weight decay"),
    "skip_layer": ("integer", 4, "add a skip connection to the output vector of every skip_layer layers."),
    "num_rgb_channels": ("integer", 3, "the number of RGB channels."),
    "num_sigma_channels": ("integer", 1, "the number of density channels."),
    "randomized": ("bool", True, "use randomized stratified sampling."),
    "min_deg_point": ("integer", 0, "Minimum degree of positional encoding for points."),
    "max_deg_point": ("integer", 10, "Maximum degree of positional encoding for points."),
    "deg_view": ("integer", 4, "Degree of positional encoding for viewdirs."),
    "num_coarse_samples": ("integer", 64, "the number of samples on each ray for the coarse model."),
    "num_fine_samples": ("integer", 128, "the number of samples on each ray for the fine model."),
    "use_viewdirs": ("bool", True, "use view directions as a condition."),
    "sh_deg": ("integer", -1, "set to use SH output up to given degree, -1 = disable."),
    "sg_dim": ("integer", -1, "set to use spherical gaussians (SG). -1 = disable"),
    "noise_std": ("float", None, "std dev of noise added to regularize sigma output."),
    "lindisp": ("bool", False, "sampling linearly in disparity rather than depth."),
    "net_activation": ("string", "relu", "activation function used within the MLP."),
    "rgb_activation": ("string", "sigmoid", "activation function used to produce RGB."),
    "sigma_activation": ("string", "relu", "activation function used to produce density."),
    "legacy_posenc_order": ("bool", False, "revert the positional encoding feature order to an older version."),
    "lr_init": ("float", 5e-4, "The initial learning rate."),
    "lr_final": ("float", 5e-6, "The final learning rate."),
    "lr_delay_steps": ("integer", 0, "steps at the beginning of training to reduce the learning rate"),
    "lr_delay_mult": ("float", 1.0, "A multiplier on the learning rate when the step is < lr_delay_steps"),
    "max_steps": ("integer", 1000000, "the number of optimization steps."),
    "save_every": ("integer", 10000, "the number of steps to save a checkpoint."),
    "print_every": ("integer", 1000, "the number of steps between reports to tensorboard."),
    "render_every": ("integer", 20000, "the number of steps to render a test image."),
    "gc_every": ("integer", 5000, "the number of steps to run python garbage collection."),
    "sparsity_weight": ("float", 1e-3, "Sparsity loss weight"),
    "sparsity_length": ("float", 0.05, "Sparsity loss 'length' for alpha calculation"),
    "sparsity_radius": ("float", 1.5, "Sparsity loss point sampling box 1/2 side length"),
    "sparsity_npoints": ("integer", 10000, "Number of samples for sparsity loss"),
    "eval_once": ("bool", True, "evaluate the model only once if true."),
    "save_output": ("bool", True, "save predicted images to disk if True."),
    "chunk": ("integer", 8192, "the size of chunks for evaluation inferences."),
    "approx_eval_skip": ("integer", 1, "Evaluates only every x images"),
    # octree side (octree/nerf/utils.py:210-225)
    "renderer_step_size": ("float", 1e-4, "step size epsilon in volume render."),
    "no_early_stop": ("bool", False, "If set, does not use early stopping in the octree renderer."),
}

_DEFINERS = {"string": flags.DEFINE_string, "bool": flags.DEFINE_bool, "integer": flags.DEFINE_integer,
             "float": flags.DEFINE_float}


def define(table):
    for name, (kind, default, helptxt) in table.items():
        if name not in FLAGS:
            _DEFINERS[kind](name, default, helptxt)


# defaults that differ on the octree side of the reference (octree/nerf/utils.py:44-219 vs nerf_sh/nerf/utils.py:60-253)
_OCTREE_DEFAULTS = {"chunk": 81920, "gc_every": 10000, "print_every": 500, "render_every": 10000, "save_every": 5000,
                    "net_activation": "ReLU", "rgb_activation": "Sigmoid", "sigma_activation": "ReLU"}


def define_flags(octree=False):
    """nerf_sh side by default; octree=True applies the octree side's defaults (octree.extraction / optimization /
    evaluation are separate programs in the reference, each with its own copy of define_flags)."""
    define(_COMMON)
    if octree:
        for name, value in _OCTREE_DEFAULTS.items():
            FLAGS.set_default(name, value)


def update_flags(args):
    """utils.update_flags (nerf_sh/nerf/utils.py:233-244): `<config>.yaml` overrides existing flags only."""
    if getattr(args, "config", None) is None:
        return
    pth = os.path.expanduser(args.config + ".yaml")
    if not os.path.exists(pth):
        from ..presets import nerf_sh_preset          # the reference's shipped configs, by base name
        configs = nerf_sh_preset(args.config)
        if configs is None:
            raise FileNotFoundError(pth)
    else:
        with open(pth, "r") as fin:
            configs = yaml.load(fin, Loader=yaml.FullLoader)
    known = set(args) if hasattr(args, "__iter__") else set(dir(args))
    invalid = sorted(set(configs.keys()) - known)
    if invalid:
        raise ValueError(f"Invalid args {invalid} in {pth}.")
    for k, v in configs.items():
        setattr(args, k, v)


def check_flags(args, require_data=True, world=1):
    """utils.check_flags (nerf_sh/nerf/utils.py:247-253)."""
    if args.train_dir is None:
        raise ValueError("train_dir must be set. None set now.")
    if require_data and args.data_dir is None:
        raise ValueError("data_dir must be set. None set now.")
    if args.batch_size % world != 0:
        raise ValueError("Batch size must be divisible by the number of devices.")


def check_scope(args):
    """features of the reference this path does not cover: fail loudly instead of training something else."""
    if args.use_viewdirs:
        raise NotImplementedError("use_viewdirs (vanilla NeRF colour head) is outside the NeRF-SH path")
    if args.sg_dim > 0:
        raise NotImplementedError("spherical gaussians (sg_dim) are outside the NeRF-SH path")
    if args.dataset not in ("blender", "llff", "nsvf"):
        raise NotImplementedError(f"dataset {args.dataset!r}: blender, llff or nsvf expected")
    if (args.net_depth, args.net_width, args.skip_layer, args.min_deg_point, args.max_deg_point) != (8, 256, 4, 0, 10):
        raise NotImplementedError("the fused kernel is built for the 8x256 trunk, skip 4, posenc degrees 0..10")
    if tuple(str(a).lower() for a in (args.net_activation, args.rgb_activation, args.sigma_activation)) != (
            "relu", "sigmoid", "relu"):
        raise NotImplementedError("activations other than relu / sigmoid / relu")   # models.py:280-281 raise the same
    if args.legacy_posenc_order:
        raise NotImplementedError("legacy_posenc_order is outside the scope of this path")
    if (args.render_path or args.spherify) and args.dataset != "llff":
        raise ValueError("render_path / spherify apply to the llff dataset only")        # datasets.py:194-195,496-497
