"""x266_b200 -- B200 (sm_100a) drop-in for the block-parallel encode hot path of chenm001/x266.

The product is the C-ABI shared library ``libx266_b200.so`` (see include/x266_b200.h).  This package
is only the thin Python binding used by tests/ and bench.py: ctypes prototypes with the reference's
own names (src_tb/dct32.c, src_tb/satd.c) and the batched ``x*`` entry points.  Nothing here computes:
if the CUDA library is missing or unusable the import of ``lib()`` raises -- there is no CPU path.
"""
from .binding import (  # noqa: F401
    LIB_PATH, lib, build, last_error, kernel_launches, set_dct_variant, tune, host_register, host_unregister, host_copy_threads,
    DCT_AUTO, DCT_BFLY, DCT_IMMA,
    partialButterfly32, satd8x8, g_t32,
    xDct32Batch, xDctNBatch, xSatd8x8Batch, xSatd8x8Search, xIntra32Pred,
    xDct32BatchDev, xDctNBatchDev, xSatd8x8BatchDev, xSatd8x8SearchDev, xIntra32PredDev, xPartialButterfly32Dev,
    xFrameResiDct32, xFrameResiDct32Dev, xConvInputFmtDev, xConvOutput420Dev, xConvInputFmt, xConvOutput420, xSatd8x8SearchTiled, xSatd8x8SearchTiledDev, xSad8x8SearchTiledDev,
    xTranspose32x32Batch, xTranspose32x32BatchDev,
    xIntra32Decide, xIntra32DecideDev, xIntra32PredModes, xIntra32PredModesDev, xIntra32EncodeBlock, xIntra32EncodeBlockDev, xIntra32Recon, xIntra32ReconDev, xQuantDequantDev, xIdct32Batch, xIdct32BatchDev, xDct32BatchMultiGpu,
    sad, xSad8x8Search, xSad8x8SearchDev, xIntra32MmaTable,
    xSad8x8SearchU16Dev, xSatd8x8SearchU16Dev, xSatd8x8SearchTiledU16Dev, xSad8x8SearchTiledU16Dev,
    bdpi_dct_block, bdpi_satd_block, X266Error,
)
