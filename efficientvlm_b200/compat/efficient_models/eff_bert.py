"""efficient_models.eff_bert -> B200 implementation."""
from efficientvlm_b200.eff_bert import *  # noqa: F401,F403
from efficientvlm_b200.eff_bert import (BertConfig, BertEmbeddings, BertEncoder, BertForMaskedLM, BertLayer, BertLMHeadModel,  # noqa: F401
                                        BertModel, BertOnlyMLMHead, BertPreTrainedModel, LabelSmoothSoftmaxCEV1)
