from .transforms import DeviceTransform, DeviceBatchCollator, build_transforms, get_size, resample_coeffs
