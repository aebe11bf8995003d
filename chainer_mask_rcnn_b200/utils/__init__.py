# flake8: noqa
from .bbox import generate_anchor_base
from .bbox import non_maximum_suppression
from .bbox import nms_suppression_bitmask
from .proposal_creator import ProposalCreator
from . import config
