{
  "targets": [{
    "target_name": "snarkb200_napi",
    "sources": ["snarkb200_napi.cc"],
    "include_dirs": ["<!@(node -p \"require('node-addon-api').include\")", "../../include"],
    "libraries": ["-L<(module_root_dir)/../../snarkjs_b200", "-lsnarkb200", "-Wl,-rpath,<(module_root_dir)/../../snarkjs_b200"],
    "defines": ["NAPI_CPP_EXCEPTIONS"],
    "cflags_cc!": ["-fno-exceptions"]
  }]
}
