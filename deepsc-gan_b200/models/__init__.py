from .modules import (CustomSchedule, Decoder, DecoderLayer, Encoder, EncoderLayer, SD, SDecoder, SE, SEncoder, STD, STE,
                      StarTransformerDecoderLayer, StarTransformerEncoderLayer, create_look_ahead_mask, create_masks,
                      create_padding_mask, loss_function, postional_encoder, sublayer1, sublayer2)
from .transceiver import (Channel_Decoder, Channel_Encoder, Channels, Transeiver, Transeiver_GAN, Transeiver_Star,
                          Transeiver_star)
from .gan import G
