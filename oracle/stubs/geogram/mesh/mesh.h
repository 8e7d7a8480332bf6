// TEST INFRASTRUCTURE ONLY.  Stand-in for geogram's <geogram/mesh/mesh.h> (geogram v1.7.5 is an un-vendored
// download of the reference, cmake/rpdDownloadExternal.cmake:17-48, absent here): just enough for the reference's
// src/IO/IO_CUDA/io_cuda.cxx to compile in place -- it only needs a 3-vector type.
#pragma once
struct Vector3 {
  double x, y, z;
  Vector3() : x(0), y(0), z(0) {}
  Vector3(double a, double b, double c) : x(a), y(b), z(c) {}
};
