"""Training path (autograd through libdod kernels).  Filled in after the inference path."""


def detector_forward_train(model, pixel_values):
    raise NotImplementedError(
        "libdod training forward/backward is not available in this build; wrap inference calls "
        "in torch.no_grad() (there is deliberately no PyTorch fallback)")
