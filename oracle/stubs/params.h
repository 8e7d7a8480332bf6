// TEST INFRASTRUCTURE ONLY.  Stand-in for src/inputs/params.h: only passed through to is_slice_by_plane.
#pragma once
struct Parameter {};
