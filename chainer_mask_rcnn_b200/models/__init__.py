# flake8: noqa
from .mask_rcnn import MaskRCNN
from .mask_rcnn_resnet import MaskRCNNResNet
from .mask_rcnn_resnet import ResNetRoIHead
from .mask_rcnn_train_chain import MaskRCNNTrainChain
from .region_proposal_network import RegionProposalNetwork
from .resnet_extractor import ResNet101Extractor
from .resnet_extractor import ResNet50Extractor
from .resnet_extractor import ResNetExtractorBase
from . import utils
