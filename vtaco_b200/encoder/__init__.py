"""Encoder registry — drop-in for reference src/encoder/__init__.py:11-20 restricted to the
conv-occupancy hot path ('pointnet_local_pool'); the other reference encoders are out of
scope."""
from . import pointnet

encoder_dict = {
    'pointnet_local_pool': pointnet.LocalPoolPointnet,
}
