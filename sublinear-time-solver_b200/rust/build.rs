fn main() {
    // libsublinear_b200.so is built by `make -C sublinear-time-solver_b200`
    let dir = std::env::var("SUBLINEAR_B200_LIB_DIR").unwrap_or_else(|_| "..".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=sublinear_b200");
}
