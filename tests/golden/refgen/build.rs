// Writes $OUT_DIR/ref_mods.rs: the reference's PRIVATE DCT modules, compiled into this program unmodified through #[path]
// (vid_dup_finder_lib/src/video_hashing/{raw_dct_ops,dct_3d}.rs are not exported by the crate; `VideoHash::from_frames`,
// the only caller, is pub(crate)).  Their `use crate::definitions::{DCT_SIZE, HASH_SIZE}` resolves against the two
// constants restated in main.rs (definitions.rs:34,36).
use std::{env, fs, path::PathBuf};

fn main() {
    let here = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let reference = env::var("VDF_REFERENCE_DIR").map(PathBuf::from).unwrap_or_else(|_| here.join("reference"));
    let reference = reference.canonicalize().expect("set VDF_REFERENCE_DIR or symlink the reference checkout as ./reference");
    let vh = reference.join("vid_dup_finder_lib/src/video_hashing");
    for f in ["raw_dct_ops.rs", "dct_3d.rs"] {
        assert!(vh.join(f).exists(), "{} not found under {}", f, vh.display());
        println!("cargo:rerun-if-changed={}", vh.join(f).display());
    }
    let out = PathBuf::from(env::var("OUT_DIR").unwrap()).join("ref_mods.rs");
    fs::write(
        &out,
        format!(
            "pub mod video_hashing {{\n    #[path = {:?}]\n    pub mod raw_dct_ops;\n    #[path = {:?}]\n    pub mod dct_3d;\n}}\n",
            vh.join("raw_dct_ops.rs"),
            vh.join("dct_3d.rs")
        ),
    )
    .unwrap();
    println!("cargo:rerun-if-env-changed=VDF_REFERENCE_DIR");
}
