"""hotformerloc_b200 -- B200-native (sm_100a) implementation of the HOTFormerLoc
embedding hot path behind the reference's own interfaces
(models/model_factory.py, HOTFormerLoc.forward(batch), config schema,
eval/pnv_evaluate.py).  See DESIGN.md."""
__version__ = '0.1.0'
