from .transforms import (DeviceBatchCollator, DeviceBatchCollatorTriplet, DeviceTransform, build_transforms, get_size,
                         resample_coeffs)
