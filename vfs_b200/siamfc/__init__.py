from .heads import SiamConvFC, SiamFC
from .tracker import DEFAULT_CFG, Net, TrackerSiamFC, build_cfg
from .train import Adam, create_labels, siamfc_loss, xcorr_backward

__all__ = ['SiamFC', 'SiamConvFC', 'Net', 'TrackerSiamFC', 'DEFAULT_CFG', 'build_cfg', 'Adam', 'create_labels',
           'siamfc_loss', 'xcorr_backward']
