from .bounding_box import BoxList, cache_source_flags, is_source_image
from .image_list import ImageList, to_image_list
