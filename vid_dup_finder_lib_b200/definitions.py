"""Constants of vid_dup_finder_lib/src/definitions.rs (same names, same values)."""
import enum

DEFAULT_SEARCH_TOLERANCE = 0.35  # definitions.rs:5
DEFAULT_VID_HASH_SKIP_FORWARD = 15.0  # definitions.rs:18
DEFAULT_VID_HASH_DURATION = 10.0  # definitions.rs:29
DCT_SIZE = 16  # definitions.rs:34
HASH_SIZE = 10  # definitions.rs:36
TOLERANCE_SCALING_FACTOR = float(HASH_SIZE**3)  # definitions.rs:40
HASH_BITS = HASH_SIZE**3  # definitions.rs:42
HASH_WORDS = -(-HASH_BITS // 64)  # definitions.rs:43 (usize::BITS = 64)


class Cropdetect(enum.Enum):
    """definitions.rs:46-54.  Letterbox is the library default; Motion is the non-default option of SURVEY.md section 8(f) N4."""

    NONE = 0
    LETTERBOX = 1
    MOTION = 2


def tolerance_to_int(tolerance: float) -> int:
    """(tolerance * TOLERANCE_SCALING_FACTOR) as u32 -- search_algorithm.rs:82; Rust's cast truncates toward
    zero, saturates, and maps NaN to 0."""
    v = float(tolerance) * TOLERANCE_SCALING_FACTOR
    if v != v or v <= 0.0:
        return 0
    if v >= 4294967295.0:
        return 4294967295
    return int(v)
