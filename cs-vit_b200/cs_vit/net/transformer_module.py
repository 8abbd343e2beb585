"""Import-path compatibility: ``cs_vit.net.transformer_module`` is where the reference defines these classes."""
from .blocks import (MHA, CrossAttnDecoder, DecoderBlock, EncoderBlock, FeedForwardNetwork,  # noqa: F401
                     PositionalEncoding)
