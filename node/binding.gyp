{
  # node-gyp rebuild  (from node/; needs libfmcw_cuda.so built first: python -m fmcw_radar_processing_b200.build)
  "targets": [
    {
      "target_name": "fmcw_napi",
      "sources": ["fmcw_napi.cc"],
      "include_dirs": ["../include"],
      "cflags_cc": ["-std=c++17"],
      "defines": ["NAPI_VERSION=8"],
      "libraries": ["-L<(module_root_dir)/../fmcw_radar_processing_b200", "-lfmcw_cuda",
                    "-Wl,-rpath,<(module_root_dir)/../fmcw_radar_processing_b200"]
    }
  ]
}
