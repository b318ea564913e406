"""Developer tool: per-parameter gradient error of the libdod training path vs oracle autograd."""
import sys
import torch
from helpers import build_product_model, detector_oracle, manifest, synth
from test_train_gpu import _oracle_grads

case = sys.argv[1] if len(sys.argv) > 1 else "c1_small_deform"
man = manifest()[case]
model, sd, kw = build_product_model(case, device="cuda", dropout=0.0)
model.train()
x = synth.make_images(man["batch"], *man["hw"], seed=man["image_seed"])
out = model(x.cuda())
(torch.nn.functional.softplus(out["pred_logits"]).sum() + ((out["pred_boxes"] - 0.3) ** 2).sum()).backward()
ref_sd, ref_out = _oracle_grads(sd, x, kw)
verbose = len(sys.argv) > 2
n_dec = kw["num_decoder_layers"]
for name, p in model.named_parameters():
    if not p.requires_grad or name.startswith("decoder.reference_points."):
        continue
    key = name.replace("layers.0.", f"layers.{n_dec - 1}.") if (kw["use_deformable"] and name.startswith("decoder.decoder.layers.0.")) else name
    ref = ref_sd[key].grad
    got = p.grad.detach().float().cpu()
    err = ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-9)).item()
    cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
    if verbose or cos < 0.995 or err > 0.1:
        print(f"{err:9.3e} cos={cos:+.4f} |ref|={ref.abs().max().item():.3e} |got|={got.abs().max().item():.3e} {name}")
