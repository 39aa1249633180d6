from .heads import SiamConvFC, SiamFC

__all__ = ['SiamFC', 'SiamConvFC']
