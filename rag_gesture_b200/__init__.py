"""B200-native guided-DDIM + exemplar-retrieval hot path of RAG-Gesture."""
__version__ = "0.1.0"
