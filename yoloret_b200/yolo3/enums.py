"""Enumerations of the reference flag surface (reference code/yolo3/enums.py:1-41)."""
from enum import Enum, unique


@unique
class MODE(Enum):
    TRAIN = 0
    IMAGE = 1
    VIDEO = 2
    TFLITE = 3
    SERVING = 4
    MAP = 5
    PRUNE = 6
    TFJS = 7
    TRAIN_BACKBONE = 8


@unique
class OPT(Enum):
    XLA = 0
    DEBUG = 1
    MKL = 2


@unique
class BACKBONE(Enum):
    MOBILENETV2x75 = 0
    MOBILENETV2x14 = 1
    EFFICIENTNETB3 = 2
    EFFICIENTNETLITE0 = 3  # derived config of BASELINE.json (not in the reference)


@unique
class BOX_LOSS(Enum):
    MSE = 0
    GIOU = 1


@unique
class DATASET_MODE(Enum):
    TRAIN = 0
    VALIDATE = 1
    TEST = 2
