from .nets import *      # noqa: F401,F403
from .nets import (Net, ImpalaEncoderProjNet, NatureEncoderProjNet, LocoTransformer,  # noqa: F401
                   Transformer, ZeroNet)
from .base import *      # noqa: F401,F403
from .base import (MLPBase, NatureEncoder, NatureFuseEncoder, TransformerEncoder,  # noqa: F401
                   LocoTransformerEncoder, RLProjection, Flatten)
from .init import *      # noqa: F401,F403
from .init import (basic_init, uniform_init, orthogonal_init, layer_init, _fanin_init,  # noqa: F401
                   _uniform_init, _constant_bias_init, _orthogonal_init)
