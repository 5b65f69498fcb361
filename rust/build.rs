// Link against the C-ABI library built by `python -m rustpde_b200.build`.
fn main() {
    let dir = std::env::var("RUSTPDE_B200_LIB_DIR").unwrap_or_else(|_| "../rustpde_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rustpde_b200");
}
