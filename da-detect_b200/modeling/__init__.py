from .detector import GeneralizedRCNN, build_detection_model
