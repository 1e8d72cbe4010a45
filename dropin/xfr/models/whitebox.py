"""`xfr.models.whitebox` served by the B200 engine: the same public names as reference python/xfr/models/whitebox.py
(WhiteboxNetwork, WhiteboxSTResnet, WhiteboxLightCNN, Whitebox_resnet50_128, Whitebox_senet50_256, Whitebox)."""
from xfr_b200.whitebox import (Whitebox, WhiteboxNetwork, WhiteboxSTResnet, WhiteboxLightCNN,  # noqa: F401
                               Whitebox_resnet50_128, Whitebox_senet50_256)

__all__ = ['Whitebox', 'WhiteboxNetwork', 'WhiteboxSTResnet', 'WhiteboxLightCNN', 'Whitebox_resnet50_128', 'Whitebox_senet50_256']
