from .cfgnode import CfgNode
from .defaults import cfg, get_cfg_defaults
