def build_resnet_backbone(cfg, input_shape):
    raise NotImplementedError("d2lite: MSRA ResNet is off the BASELINE path (RESNETS.TORCHVISION True, config.py:141)")
