"""GPU: CUDA-graph replay of the forward returns exactly what the eager launch sequence returns."""
import pytest
import torch

from helpers import build_product_model, synth

pytestmark = pytest.mark.gpu


def test_graphed_forward_matches_eager():
    from dino_detector.runtime import GraphedDetector
    model, sd, kw = build_product_model("c1_small_deform", device="cuda")
    x1 = synth.make_images(2, 224, 224, seed=1).cuda()
    x2 = synth.make_images(2, 224, 224, seed=2).cuda()
    graphed = GraphedDetector(model, x1)
    for x in (x1, x2, x1):
        with torch.no_grad():
            eager = {k: v.clone() for k, v in model(x).items()}
        out = graphed(x)
        torch.cuda.synchronize()
        for k in eager:
            assert torch.equal(out[k], eager[k]), k
    with pytest.raises(ValueError):
        graphed(torch.rand(1, 3, 224, 224).cuda())
