{
  "targets": [{
    "target_name": "carta1_b200",
    "sources": ["carta1_napi.c"],
    "defines": ["CARTA1_NAPI_USE_NODE_HEADERS", "NAPI_VERSION=8"],
    "include_dirs": ["../../include"],
    "libraries": ["-L<(module_root_dir)/..", "-lcarta1_b200", "-Wl,-rpath,<(module_root_dir)/.."]
  }]
}
