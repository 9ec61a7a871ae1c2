"""models.xbert -> B200 implementation (gates default to None)."""
from efficientvlm_b200.eff_bert import *  # noqa: F401,F403
from efficientvlm_b200.eff_bert import (BertConfig, BertEmbeddings, BertEncoder, BertForMaskedLM, BertLayer, BertLMHeadModel,  # noqa: F401
                                        BertModel, BertOnlyMLMHead, BertPreTrainedModel)
