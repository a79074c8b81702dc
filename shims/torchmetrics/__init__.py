"""Stand-in for the two torchmetrics classes reference train.py uses (train.py:11,72-73,101-104,114).
SSIM runs on our fused CUDA kernels (3d-gaussian-splatting_b200/loss.py -> csrc/loss.cu)."""
import torch

from .functional import peak_signal_noise_ratio, structural_similarity_index_measure

__version__ = "0.0-gs-b200-shim"


class _Metric:
    def to(self, *_a, **_k):
        return self

    def cuda(self, *_a, **_k):
        return self

    def eval(self):
        return self

    def reset(self):
        pass

    def __call__(self, preds, target):
        return self.forward(preds, target)


class StructuralSimilarityIndexMeasure(_Metric):
    def __init__(self, gaussian_kernel=True, sigma=1.5, kernel_size=11, reduction="elementwise_mean", data_range=None,
                 k1=0.01, k2=0.03, **_kw):
        if not (gaussian_kernel and sigma == 1.5 and kernel_size == 11 and k1 == 0.01 and k2 == 0.03
                and reduction == "elementwise_mean" and data_range in (1.0, 1, (0.0, 1.0))):
            raise NotImplementedError("shim: only the configuration of reference train.py:72 "
                                      "(defaults, data_range=1.0) is implemented")

    def forward(self, preds, target):
        return structural_similarity_index_measure(preds, target, data_range=1.0)


class PeakSignalNoiseRatio(_Metric):
    def __init__(self, data_range=None, **_kw):
        self.data_range = data_range

    def forward(self, preds, target):
        return peak_signal_noise_ratio(preds, target, data_range=self.data_range)
