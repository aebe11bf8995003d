# flake8: noqa
from .bbox_tools import bbox2loc
from .bbox_tools import bbox_iou
from .bbox_tools import loc2bbox
from .anchor_target_creator import AnchorTargetCreator
from .proposal_target_creator import ProposalTargetCreator
from .device_targets import DeviceAnchorTargetCreator
from .device_targets import DeviceProposalTargetCreator
from .device_targets import GroundTruth
from .device_targets import PackedMasks
