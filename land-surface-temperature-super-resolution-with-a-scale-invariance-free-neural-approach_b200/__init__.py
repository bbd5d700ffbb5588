"""B200-native SIF-NN-SR ModelB hot path (inference, SR1/SR2 training step, Adam, data
parallelism).  Host side: Python/PyTorch for device memory, streams and
``torch.distributed``; arithmetic: hand-written sm_100a CUDA in ``libsifnn_b200.so``
behind the C-ABI of ``include/sifnn.h``.  No Triton, no multi-backend dispatch, no CPU
fallback.

The directory name is fixed by the build contract and is not a Python identifier; import
it as ``sifnn_b200`` (see ``sifnn_b200.py`` at the repository root), or use the
repository-root ``model.py`` for reference-compatible ``from model import ModelB_2``.
"""
from ._lib import SifnnError, build, load, LIB_PATH  # noqa: F401
from .model import (ModelB_2, DoubleConvolution, UpBlock, ResidualConnection, DownBlock_pool, DownBlock,  # noqa: F401
                    ResBridgeBlock, Serf, activation_functions, bicubic4_cat)
from .losses import sr_losses, sr1_losses, sr2_losses, loss_fwd_bwd  # noqa: F401
from .trainer import Trainer  # noqa: F401
from .parallel import block_partition, shard_batch, BucketedAllReduce  # noqa: F401
from .tile import super_resolve_tile, super_resolve_tile_host, super_resolve_geotiff, window_list, owned_rows  # noqa: F401
from .serving import PipelinedInference  # noqa: F401
from .fit import model_checkpoint, fit, train_epoch, eval_epoch, save_model, load_model, save_metrics  # noqa: F401
from .dataset import ModisDatasetB, PinnedBatchLoader, read_geotiff, save_geotiff, upsampling  # noqa: F401



def set_tensor_cores(on: bool) -> None:
    """Allow (default) or forbid the tcgen05 kernels in ModelB_2 / Trainer.  ``False`` = strict fp32 SIMT everywhere."""
    load().sifnn_set_tensor_cores(1 if on else 0)


def tensor_cores_enabled() -> bool:
    return bool(load().sifnn_get_tensor_cores())


__version__ = "0.1.0"
