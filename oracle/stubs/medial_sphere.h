// TEST INFRASTRUCTURE ONLY.  Stand-in for src/matbase/medial_sphere.h (needs geogram): save_convex_cells_houdini
// takes the sphere vector and never reads it.
#pragma once
struct MedialSphere {};
