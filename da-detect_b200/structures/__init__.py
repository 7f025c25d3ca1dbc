from .bounding_box import BoxList
from .image_list import ImageList, to_image_list
