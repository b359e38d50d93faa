"""tiny-newsrec_b200: B200-native (sm_100a) implementation of the Tiny-NewsRec
training / scoring hot path behind the reference's own module API.

Layout
  csrc/           hand-written CUDA kernels + the C-ABI (include/tinyrec.h)
  _lib.py         ctypes binding of libtinyrec.so (fails loudly if missing)
  ops.py          tensor-level wrappers over the C-ABI entry points
  engine.py       workspace + forward/backward orchestration of the encoder path
  model_bert.py   drop-in for Tiny-NewsRec/model_bert.py (Model, ModelBert, ...)
  model_bert_2.py drop-in for Tiny-NewsRec/model_bert_2.py (teacher ModelBert)
  optim.py        fused Adam(amsgrad) + NCCL data-parallel optimizer wrapper
  dataloader.py   device-resident batch assembly (index gathers)
  run.py          train / test / get_teacher_emb drivers
  synth.py        deterministic random-init weights + synthetic MIND-shaped data

Import as ``import tinyrec`` (see /tinyrec.py).  There is no CPU fallback.
"""
__version__ = "0.1.0"
