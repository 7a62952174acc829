"""B200-native fused dual-corpus retrieval engine: drop-in for the search path of
ClipABit/Multimodal-Audio-Search (`DualPipelineAudioSearch.search_with_fusion`)."""
from .engine import (DualPipelineAudioSearch, PipelineStats, accelerate,  # noqa: F401
                     load_library, save_library)
from .segment_table import SegmentRecord, SegmentTable  # noqa: F401
from .batcher import SearchBatcher  # noqa: F401
from .index import SearchResult, SegmentIndex, synth_queries  # noqa: F401
from .query_weights import analyze_query_for_weights  # noqa: F401
from .sharded import ShardedSearcher, shard_range  # noqa: F401

__all__ = ["DualPipelineAudioSearch", "PipelineStats", "accelerate", "SegmentIndex", "SearchResult",
           "synth_queries", "analyze_query_for_weights", "ShardedSearcher", "shard_range",
           "SegmentTable", "SegmentRecord", "save_library", "load_library", "SearchBatcher"]
