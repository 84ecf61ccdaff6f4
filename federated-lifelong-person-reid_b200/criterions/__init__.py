from .losses import CriterionModule, CrossEntropyLabelSmooth, TripletLoss, DistillKL  # noqa: F401
from .losses import euclidean_dist, cosine_dist, kl_distance  # noqa: F401

criterions = {
    "cross_entropy": CrossEntropyLabelSmooth,
    "triplet_loss": TripletLoss,
    "kd_loss": DistillKL,
}
