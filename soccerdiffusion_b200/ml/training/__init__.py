from soccerdiffusion_b200.ml.training.data import DevicePrefetcher, bind_to_numa_node, gpu_numa_node  # noqa: F401
from soccerdiffusion_b200.ml.training.graph import GraphedTrainStep  # noqa: F401
from soccerdiffusion_b200.ml.training.optim import FusedAdamW  # noqa: F401
from soccerdiffusion_b200.ml.training.step import (  # noqa: F401
    BucketedAllReduce,
    allreduce_gradients,
    broadcast_parameters,
    distill_step,
    train_step,
)
