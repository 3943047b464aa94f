class LearnedPerceptualImagePatchSimilarity:
    """eval.py:153 asks for LPIPS with an AlexNet trunk (PeopleSnapshot evaluation): its weights are not available offline
    and that network is not part of this package (the training loss is LPIPS-VGG: gomavatar_b200.lpips)."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("torchmetrics stand-in: LPIPS(net_type='alex') is not provided; install torchmetrics "
                                  "for Evaluator_snapshot")
