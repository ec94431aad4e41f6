fn main() {
    // libg16b200.so is built by `make -C crescent_credentials_b200/csrc`; point G16_B200_LIB_DIR at its directory.
    let dir = std::env::var("G16_B200_LIB_DIR").unwrap_or_else(|_| "/usr/local/lib".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=g16b200");
}
