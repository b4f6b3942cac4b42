"""ConvolutionalOccupancyNetwork container and decoder registry — drop-in for
reference src/conv_onet/models/__init__.py:7-197."""
import torch
import torch.nn as nn
from torch import distributions as dist

from . import decoder

# Decoder dictionary (reference models/__init__.py:7-12).  Only 'simple_local' is on the
# conv-occupancy hot path; the other reference decoders are out of scope (SURVEY §2 rows 9,10).
decoder_dict = {
    'simple_local': decoder.LocalDecoder,
}


def _bernoulli(logits):
    # The reference builds dist.Bernoulli(logits=...) with argument validation on, which
    # costs a host sync per call; callers only read .logits / .probs.
    return dist.Bernoulli(logits=logits, validate_args=False)


class ConvolutionalOccupancyNetwork(nn.Module):
    ''' Occupancy Network class (reference models/__init__.py:15-197).

    Args:
        decoder (nn.Module): decoder network
        encoder (nn.Module): encoder network
        device (device): torch device
    '''

    def __init__(self, decoder, encoder=None, encoder_hand=None, encoder_img=None, encoder_t2d=None, device=None):
        super().__init__()
        self.decoder = decoder.to(device) if decoder is not None else None
        self.encoder = encoder.to(device) if encoder is not None else None
        self.encoder_hand = encoder_hand.to(device) if encoder_hand is not None else None
        self.encoder_img = encoder_img.to(device) if encoder_img is not None else None
        self.encoder_t2d = encoder_t2d.to(device) if encoder_t2d is not None else None
        self._device = device

    def forward(self, p, inputs, imgs=None, sample=True, **kwargs):
        c = self.encode_inputs(inputs)
        self.encode_hand_inputs(inputs)
        return self.decode(p, c, **kwargs)

    def encode_inputs(self, inputs):
        if self.encoder is not None:
            return self.encoder(inputs)
        return torch.empty(inputs.size(0), 0)

    def encode_hand_inputs(self, inputs):
        if self.encoder_hand is not None:
            return self.encoder_hand(inputs)
        return torch.empty(inputs.size(0), 0)

    def encode_hand_mano(self, inputs):
        return self.encoder_hand.forward_mano(inputs)

    def encode_img_inputs(self, imgs):
        if self.encoder_img is not None:
            B, F, C, H, W = imgs.size()
            c_list = []
            for b_idx in range(B):
                imgs_in = imgs[b_idx].reshape(F, C, H, W)
                c_list.append(self.encoder_img(imgs_in).reshape(1, F, -1))
            return torch.cat(c_list, dim=0)
        return torch.empty(imgs.size(0), 0)

    def encode_t2d(self, inputs, imgs):
        pred_depth = self.encoder_t2d.encode_img_inputs(imgs)
        c_hand = self.encoder_t2d.encode_hand_inputs(inputs)
        return pred_depth, c_hand

    def decode(self, p, c, **kwargs):
        return _bernoulli(self.decoder(p, c, **kwargs))

    def decode_img(self, p, c, c_img=None, **kwargs):
        return _bernoulli(self.decoder.forward_img(p, c, c_img, **kwargs))

    def decode_contact(self, p, c, **kwargs):
        logits, pred_contact = self.decoder.forward_contact(p, c, **kwargs)
        return _bernoulli(logits), pred_contact

    def to(self, device):
        model = super().to(device)
        model._device = device
        return model
