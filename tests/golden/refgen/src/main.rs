//! vdf_refgen -- reference vectors for the hashing path and the hash-cache file, produced by the REAL code:
//!
//!   crop      vid_dup_finder_common::video_frames_gray::cropdetect_letterbox        (video_frames_gray.rs:201-210)
//!   cropped   VdfFrameExt::cropped(crop).to_image()                                 (video_hash_builder.rs:198-201)
//!   small     vid_dup_finder_common::crop_resize_buf(frame, 16, 16, no_crop)        (video_hash.rs:57-59, resize_gray.rs:11-54
//!                                                                                    -> fast_image_resize, Lanczos3, U8)
//!   coefs     Dct3d::from_images(small frames) / hash_bits()                        (dct_3d.rs:15-66, raw_dct_ops.rs:107-142 -> rustdct)
//!   words     BitArray<[usize; 16], Lsb0> filled bit by bit                         (video_hash.rs:63-70)
//!   cache     bincode::serde::encode_into_std_write(&HashMap<PathBuf, MtimeCacheEntry<Result<VideoHash, Error>>>,
//!             .., bincode::config::standard())                                      (base_fs_cache.rs:26,106-112)
//!
//! Everything after "crop" is exactly what VideoHash::from_frames does (it is pub(crate), so its four statements are
//! restated here around the real callee functions; the DCT modules are private and are compiled in unmodified, build.rs).
//!
//! usage: vdf_refgen <inputs.bin> <reference_vectors.json> <reference_cache.bin>
use std::{
    collections::HashMap,
    fmt::Write as _,
    fs,
    io::{BufWriter, Read, Write},
    num::NonZeroU32,
    path::PathBuf,
    time::{Duration, SystemTime},
};

use bitvec::prelude::*;
use image::GrayImage;
use rand::{rngs::StdRng, SeedableRng};
use serde::{Deserialize, Serialize};
use sha2::{Digest, Sha256};
use vid_dup_finder_common::{
    crop_resize_buf,
    video_frames_gray::{cropdetect_letterbox, VdfFrameExt},
    Crop,
};
use vid_dup_finder_lib::{Error as VdfError, VideoHash};

/// definitions.rs:34,36 -- what `crate::definitions` resolves to inside the included DCT modules
pub mod definitions {
    pub const DCT_SIZE: u32 = 16;
    pub const HASH_SIZE: u32 = 10;
}
include!(concat!(env!("OUT_DIR"), "/ref_mods.rs"));
use video_hashing::dct_3d::Dct3d;

const HASH_WORDS: usize = 16; // definitions.rs:43 on a 64-bit target
const HASH_BITS: usize = 1000; // definitions.rs:42

struct Item {
    kind: u32,
    name: String,
    w: u32,
    h: u32,
    frames: Vec<Vec<u8>>,
}

fn rd_u32(r: &mut impl Read) -> u32 {
    let mut b = [0u8; 4];
    r.read_exact(&mut b).unwrap();
    u32::from_le_bytes(b)
}

fn read_items(path: &str) -> Vec<Item> {
    let mut r = std::io::BufReader::new(fs::File::open(path).expect("inputs.bin (python tests/golden/make_reference_inputs.py)"));
    let mut magic = [0u8; 4];
    r.read_exact(&mut magic).unwrap();
    assert_eq!(&magic, b"VDFI");
    let n = rd_u32(&mut r);
    (0..n)
        .map(|_| {
            let kind = rd_u32(&mut r);
            let mut name = vec![0u8; rd_u32(&mut r) as usize];
            r.read_exact(&mut name).unwrap();
            let (w, h, nf) = (rd_u32(&mut r), rd_u32(&mut r), rd_u32(&mut r));
            let frames = (0..nf)
                .map(|_| {
                    let mut f = vec![0u8; (w * h) as usize];
                    r.read_exact(&mut f).unwrap();
                    f
                })
                .collect();
            Item { kind, name: String::from_utf8(name).unwrap(), w, h, frames }
        })
        .collect()
}

fn hex(b: &[u8]) -> String {
    let mut s = String::with_capacity(b.len() * 2);
    for x in b {
        write!(s, "{x:02x}").unwrap();
    }
    s
}

fn sha(frames: &[Vec<u8>]) -> String {
    let mut h = Sha256::new();
    for f in frames {
        h.update(f);
    }
    hex(&h.finalize())
}

/// video_hash.rs:45-73 with the real callees; returns (16^3 u8 cube [t][row][col], 1000 sign bits, 16 words)
fn from_frames(frames: Vec<GrayImage>) -> Option<(Vec<u8>, Vec<bool>, [usize; HASH_WORDS])> {
    let dct_size = NonZeroU32::new(definitions::DCT_SIZE).unwrap();
    let first = frames.first()?;
    let no_crop = Crop::from_edge_offsets((first.width(), first.height()), 0, 0, 0, 0); // :55
    let small: Vec<GrayImage> = frames.iter().map(|f| crop_resize_buf(f, dct_size, dct_size, no_crop)).collect(); // :57-59
    let cube: Vec<u8> = small.iter().flat_map(|im| im.as_raw().iter().copied()).collect();
    let dct = Dct3d::from_images(small)?; // :61
    let bits: Vec<bool> = dct.hash_bits().collect();
    assert_eq!(bits.len(), HASH_BITS);
    let mut bitarr: BitArray<[usize; HASH_WORDS], Lsb0> = BitArray::ZERO; // :63-68
    for (mut v, b) in bitarr.iter_mut().zip(bits.iter()) {
        *v = *b;
    }
    Some((cube, bits, bitarr.into_inner()))
}

fn pack_bits(bits: &[bool]) -> Vec<u8> {
    let mut out = vec![0u8; (bits.len() + 7) / 8];
    for (i, b) in bits.iter().enumerate() {
        if *b {
            out[i / 8] |= 1 << (i % 8);
        }
    }
    out
}

/// base_fs_cache.rs / processing_fs_cache.rs:23-27: the value type of the app's cache map
#[derive(Serialize, Deserialize, Clone)]
struct MtimeCacheEntry<T> {
    cache_mtime: SystemTime,
    value: T,
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    assert!(args.len() == 4, "usage: vdf_refgen <inputs.bin> <reference_vectors.json> <reference_cache.bin>");
    let items = read_items(&args[1]);
    let mut json = String::new();
    let lock = fs::read_to_string(concat!(env!("CARGO_MANIFEST_DIR"), "/Cargo.lock")).unwrap_or_default();
    let ver = |name: &str| -> String {
        let key = format!("name = \"{name}\"\nversion = \"");
        lock.find(&key).map(|p| lock[p + key.len()..].split('"').next().unwrap().to_string()).unwrap_or_else(|| "?".into())
    };
    write!(
        json,
        "{{\n\"generator\": \"vdf_refgen 0.1.0\",\n\"crates\": {{\"fast_image_resize\": \"{}\", \"rustdct\": \"{}\", \"image\": \"{}\", \"ndarray\": \"{}\", \"bitvec\": \"{}\", \"bincode\": \"{}\"}},\n\"items\": [\n",
        ver("fast_image_resize"), ver("rustdct"), ver("image"), ver("ndarray"), ver("bitvec"), ver("bincode")
    )
    .unwrap();
    let mut first = true;
    for it in &items {
        if !first {
            json.push_str(",\n");
        }
        first = false;
        write!(json, "{{\"name\": {:?}, \"kind\": {}, \"w\": {}, \"h\": {}, \"n_frames\": {}, \"input_sha256\": \"{}\"", it.name, it.kind, it.w, it.h, it.frames.len(), sha(&it.frames)).unwrap();
        match it.kind {
            0 => {
                // a stack: crop_video_frames + from_frames (video_hash_builder.rs:188-204, video_hash.rs:45-73)
                let frames: Vec<GrayImage> = it.frames.iter().map(|f| GrayImage::from_raw(it.w, it.h, f.clone()).unwrap()).collect();
                let crop = cropdetect_letterbox(&frames).expect("16 frames");
                let cropped: Vec<GrayImage> = frames.iter().map(|f| f.cropped(crop).to_image()).collect();
                let (cube, bits, words) = from_frames(cropped).expect("16 frames");
                write!(json, ", \"crop\": [{}, {}, {}, {}], \"small\": \"{}\", \"coef_positive\": \"{}\", \"hash\": [", crop.left, crop.right, crop.top, crop.bottom, hex(&cube), hex(&pack_bits(&bits))).unwrap();
                for (k, w) in words.iter().enumerate() {
                    write!(json, "{}\"{:#018x}\"", if k > 0 { ", " } else { "" }, *w as u64).unwrap();
                }
                json.push(']');
            }
            1 => {
                // one frame: crop_resize_buf alone (resize_gray.rs:11-54)
                let f = GrayImage::from_raw(it.w, it.h, it.frames[0].clone()).unwrap();
                let d = NonZeroU32::new(16).unwrap();
                let out = crop_resize_buf(&f, d, d, Crop::from_edge_offsets((it.w, it.h), 0, 0, 0, 0));
                write!(json, ", \"small\": \"{}\"", hex(out.as_raw())).unwrap();
            }
            _ => {
                // a 16^3 cube: Dct3d::from_images + hash_bits alone (dct_3d.rs:15-66)
                let frames: Vec<GrayImage> = it.frames.iter().map(|f| GrayImage::from_raw(16, 16, f.clone()).unwrap()).collect();
                let dct = Dct3d::from_images(frames).expect("16 frames");
                let bits: Vec<bool> = dct.hash_bits().collect();
                write!(json, ", \"coef_positive\": \"{}\"", hex(&pack_bits(&bits))).unwrap();
            }
        }
        json.push('}');
    }
    json.push_str("\n],\n");

    // ---- the cache file: a few Ok and Err entries, written exactly as BaseFsCache::save does (base_fs_cache.rs:106-112)
    let mut rng = StdRng::seed_from_u64(0xB200);
    let mut map: HashMap<PathBuf, MtimeCacheEntry<Result<VideoHash, VdfError>>> = HashMap::new();
    let mut entries = String::new();
    let mk_time = |s: u64, n: u32| SystemTime::UNIX_EPOCH + Duration::new(s, n);
    for k in 0..6u32 {
        let path = PathBuf::from(format!("/videos/dir {k}/clip-{k:03}.mp4"));
        let vh = VideoHash::random_hash(&mut rng).with_src_path(&path).with_duration(100 * k + 7);
        let words: Vec<String> = vh.raw_hash().collect::<Vec<bool>>().chunks(64).map(|c| {
            let mut w = 0u64;
            for (i, b) in c.iter().enumerate() {
                if *b {
                    w |= 1 << i;
                }
            }
            format!("\"{w:#018x}\"")
        }).collect();
        write!(entries, "{}{{\"key\": {:?}, \"kind\": 0, \"duration\": {}, \"mtime\": [{}, {}], \"hash\": [{}]}}", if k > 0 { ",\n" } else { "" }, path.to_str().unwrap(), 100 * k + 7, 1_700_000_000u64 + k as u64, 1000 * k, words.join(", ")).unwrap();
        map.insert(path, MtimeCacheEntry { cache_mtime: mk_time(1_700_000_000 + k as u64, 1000 * k), value: Ok(vh) });
    }
    let errs: [(&str, VdfError, i32, &str); 3] = [
        ("/videos/not a video.txt", VdfError::NotVideo, 1, ""),
        ("/videos/broken.mkv", VdfError::VidProc("frames not all same size: Expected (640, 360), Actual (320, 180)".to_string()), 2, "frames not all same size: Expected (640, 360), Actual (320, 180)"),
        ("/videos/short.mp4", VdfError::NotEnoughFrames, 3, ""),
    ];
    for (k, (p, e, kind, msg)) in errs.into_iter().enumerate() {
        write!(entries, ",\n{{\"key\": {:?}, \"kind\": {}, \"msg\": {:?}, \"mtime\": [{}, {}]}}", p, kind, msg, 1_600_000_000u64 + k as u64, 5).unwrap();
        map.insert(PathBuf::from(p), MtimeCacheEntry { cache_mtime: mk_time(1_600_000_000 + k as u64, 5), value: Err(e) });
    }
    let mut w = BufWriter::new(fs::File::create(&args[3]).unwrap());
    bincode::serde::encode_into_std_write(&map, &mut w, bincode::config::standard()).unwrap();
    w.flush().unwrap();
    write!(json, "\"cache_entries\": [\n{entries}\n]\n}}\n").unwrap();
    fs::write(&args[2], json).unwrap();
    eprintln!("{} items -> {}, cache file -> {}", items.len(), args[2], args[3]);
}
