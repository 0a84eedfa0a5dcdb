"""jspsr_b200 - B200 (sm_100a) implementation of JSPSR's spatial-propagation
refinement step behind the reference's own nn.Module interface.

    from jspsr_b200 import PostProcessor, Post_process_deconv, NLSPN

are drop-in replacements for models.components.spn.PostProcessor,
models.LRRU.Post_process_deconv and models.components.nlspn.NLSPN of
xandercai/JSPSR (same constructor, forward signature and state_dict keys).
"""
from .modules import NLSPN, Post_process_deconv, PostProcessor, generator_postprocess  # noqa: F401
from . import functional  # noqa: F401
from . import epilogue, tiles  # noqa: F401
from .epilogue import MeterRMSE, MultiLoss  # noqa: F401

__version__ = "1.0"
