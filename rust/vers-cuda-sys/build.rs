// Compiles the CUDA device layer for B200 (sm_100a) only: no fat binary, no fallback architecture, no CPU path.
fn main() {
    let root = std::path::PathBuf::from(env!("CARGO_MANIFEST_DIR")).join("../..");
    let csrc = root.join("vers_b200/csrc");
    let mut b = cc::Build::new();
    b.cuda(true)
        .cudart("static")
        .flag("-gencode")
        .flag("arch=compute_100a,code=sm_100a")
        .flag("-O3")
        .flag("-std=c++17")
        .flag("-lineinfo")
        .flag("-fmad=false") // exact-order contract (vers/src/indexes/base.rs:91-93, 119-126): never contract a*b+c
        .include(root.join("include"));
    for f in ["api.cu", "flat.cu", "kmeans.cu", "ivf.cu", "lsh.cu", "lsh_forest.cu", "peer.cu"] {
        b.file(csrc.join(f));
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    b.compile("vers_b200");
}
