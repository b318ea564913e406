"""Constructor / loss defaults of the detector (same names and values as the reference
dino_detector/config.py:1-57; callers import these by name, train.py:28-37)."""

# Training configuration
learning_rate = 1e-4
num_epochs = 50
batch_size = 8
num_workers = 4

# Debug / overfit mode
debug_mode = False
debug_dataset_size = 32
debug_epochs = 100
debug_learning_rate = 5e-4

# Distributed training
distributed_backend = "nccl"
find_unused_parameters = True

# Model configuration
dino_model_name = "facebook/dinov2-base"
lora_r = 2
lora_alpha = 1.0
hidden_dim = 768
num_queries = 50
num_decoder_layers = 3
nheads = 8
num_classes = 91
dim_feedforward = 1024
dropout = 0.1
use_deformable = True
n_points = 2
deformable_modulation = False

# Optimiser
weight_decay = 1e-4
gradient_accumulation_steps = 1
gradient_clip_val = 1.0

# Hungarian matcher
set_cost_class = 1.0
set_cost_bbox = 5.0
set_cost_giou = 2.0

# Criterion
focal_alpha = 0.25
focal_gamma = 2.0
loss_weights = {"loss_ce": 1.0, "loss_bbox": 5.0, "loss_giou": 2.0}

# ---- libdod specific (not in the reference) -------------------------------
# "bf16": bf16 operands, fp32 accumulation, fp32 residual stream (BASELINE configs 2-5).
# "fp32": 3-term bf16 split GEMMs + fp32 attention, ~1e-6 relative to an fp32 reference
#         (BASELINE config 1 parity).  Overridable per module (`model.precision = ...`)
#         or with the DOD_PRECISION environment variable.
precision = "bf16"
# reproduce the reference's result-changing quirks: matcher uses the
# predictions of image 0 for every image (matching.py:102).
reference_compat = True
