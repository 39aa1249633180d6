from .resnet import BasicBlock, Bottleneck, ResNet, make_res_layer

__all__ = ['ResNet', 'BasicBlock', 'Bottleneck', 'make_res_layer']
