{
  "targets": [{
    "target_name": "zles",
    "sources": ["addon.c"],
    "include_dirs": ["../../include"],
    "libraries": ["-L<(module_root_dir)/..", "-lzles", "-Wl,-rpath,<(module_root_dir)/.."]
  }]
}
