// TEST INFRASTRUCTURE shim: libpng is absent; the reference's Align4.cpp only touches PngImage in
// debug-only code paths (debug=false on the hot path). Just enough for PngImage.hpp to parse.
#pragma once
typedef unsigned char png_byte;
