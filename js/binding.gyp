{
  "targets": [{
    "target_name": "headtrackr_b200_addon",
    "sources": ["addon.cc"],
    "include_dirs": ["../include"],
    "libraries": ["-L<(module_root_dir)/../headtrackr_b200", "-lheadtrackr_b200",
                  "-Wl,-rpath,<(module_root_dir)/../headtrackr_b200"]
  }]
}
