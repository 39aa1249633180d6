from .heads import SiamConvFC, SiamFC
from .tracker import DEFAULT_CFG, Net, TrackerSiamFC, build_cfg

__all__ = ['SiamFC', 'SiamConvFC', 'Net', 'TrackerSiamFC', 'DEFAULT_CFG', 'build_cfg']
