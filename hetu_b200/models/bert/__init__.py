from .bert_model import (BertConfig, BertModel, BertForPreTraining, BertForMaskedLM, BertForSequenceClassification,  # noqa: F401
                         convert_bert_hf_to_ht)
