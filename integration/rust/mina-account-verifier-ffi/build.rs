// Links the B200 library.  MINA_B200_LIB_DIR = directory holding libmina_b200.so (built by
// `python -c 'import __graft_entry__ as g; g.build()'` -> mina_bridge_b200/lib).
fn main() {
    let dir = std::env::var("MINA_B200_LIB_DIR").expect("set MINA_B200_LIB_DIR to mina_bridge_b200/lib");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=mina_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
}
