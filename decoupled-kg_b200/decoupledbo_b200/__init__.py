"""decoupledbo_b200: B200-native drop-in for the discrete knowledge-gradient hot path of
quasirandom/decoupled-kg (``DiscreteKnowledgeGradient.forward`` + backward, as driven by
``DiscreteKgOptimisationSpec``).  CUDA (sm_100a) behind a C-ABI library; no CPU fallback."""

__version__ = "0.1.0"
