"""Model classes under the reference's names (chainer_mask_rcnn.models)."""
from . import utils
from .mask_rcnn import MaskRCNN
from .mask_rcnn_resnet import MaskRCNNResNet, ResNetRoIHead
from .mask_rcnn_train_chain import MaskRCNNTrainChain
from .region_proposal_network import RegionProposalNetwork
from .resnet_extractor import ResNet50Extractor, ResNet101Extractor, ResNetExtractorBase

__all__ = ['MaskRCNN', 'MaskRCNNResNet', 'ResNetRoIHead', 'MaskRCNNTrainChain',
           'RegionProposalNetwork', 'ResNet50Extractor', 'ResNet101Extractor',
           'ResNetExtractorBase', 'utils']
