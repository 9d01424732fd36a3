"""Effective ``model`` nodes of the reference's task patches, as plain dicts for tests / bench / scripts
(/root/reference/configs/_global_patches/brats.yaml:10-19, .../hecktor21.yaml:10-19,
/root/reference/configs/model/unet.yaml + src/models/unet.py:27-48 for the bare defaults)."""

BRATS_MODEL_CFG = dict(in_channels=4, num_classes=3, spatial_dims=3, channels=[32, 64, 128, 256, 512],
                       strides=[2, 2, 2, 2], num_res_units=2, norm="INSTANCE", act="RELU", dropout=0.0)
HECKTOR_MODEL_CFG = dict(in_channels=2, num_classes=1, spatial_dims=3, channels=[32, 64, 128, 256, 512],
                         strides=[2, 2, 2, 2], num_res_units=2, norm="INSTANCE", act="RELU", dropout=0.0)
BARE_DEFAULT_MODEL_CFG = dict(name="unet", num_classes=1)
