// TEST INFRASTRUCTURE ONLY.  Stand-in for src/IO/IO_CXX/io.h (needs geogram): the one helper of it that
// save_convex_cells_houdini calls (create_dir and get_timestamp are the reference's own, include/common_cxx.h).
#pragma once
inline bool is_slice_by_plane(const Vector3&, const Parameter&) { return false; }
