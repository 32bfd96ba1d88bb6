"""vid_dup_finder_lib_b200 -- B200-native (sm_100a) hot paths of Farmadupe/vid_dup_finder_lib behind the crate's
public API (vid_dup_finder_lib/src/lib.rs:132-140): VideoHash creation from decoded frame stacks, `search`,
`search_with_references`, `MatchGroup`, the tolerance option.  All bulk compute runs in hand-written CUDA
kernels through the C ABI in include/vdf_b200.h; there is no CPU fallback."""
from ._ffi import Context, Table, VdfError, default_context
from .crop import Crop
from .definitions import (DEFAULT_SEARCH_TOLERANCE, DEFAULT_VID_HASH_DURATION, DEFAULT_VID_HASH_SKIP_FORWARD,
                          TOLERANCE_SCALING_FACTOR, Cropdetect)
from .hash_cache import CacheMetadata, HashCache, load_hash_cache, save_hash_cache
from .match_group import MatchGroup, MatchGroups, TooFewEntries
from .pipeline import HashPipeline
from .search import search, search_with_references
from .video_hash import HashTable, VideoHash
from .video_hash_builder import (CreationOptions, Error, NotEnoughFrames, NotVideo, VideoHashBuilder, VidProc)

__all__ = [
    "VideoHash", "VideoHashBuilder", "CreationOptions", "search", "search_with_references", "MatchGroup", "Error",
    "NotVideo", "VidProc", "NotEnoughFrames", "Cropdetect", "DEFAULT_SEARCH_TOLERANCE", "DEFAULT_VID_HASH_DURATION",
    "DEFAULT_VID_HASH_SKIP_FORWARD", "TOLERANCE_SCALING_FACTOR", "HashTable", "Crop", "Context", "VdfError",
    "default_context", "TooFewEntries", "MatchGroups", "Table", "HashCache", "CacheMetadata", "load_hash_cache", "save_hash_cache", "HashPipeline",
]
