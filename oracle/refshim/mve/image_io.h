// refshim: MVE mve/image_io.h stand-in: "files" are images registered in memory by the glue
// (oracle/ref_glue.cpp); nothing is decoded or written (see ../README.md).
#pragma once
#include <map>
#include <string>
#include "mve/image.h"
#include "util/exception.h"

namespace mve { namespace image {

struct ImageHeaders { int width; int height; int channels; ImageType type; };

std::map<std::string, ByteImage::Ptr>& refshim_registry();   // defined in ref_glue.cpp

inline ByteImage::Ptr load_file(std::string const& filename) {
    std::map<std::string, ByteImage::Ptr>::iterator it = refshim_registry().find(filename);
    if (it == refshim_registry().end()) throw util::FileException(filename, "not registered");
    return it->second->duplicate();
}
inline ImageHeaders load_file_headers(std::string const& filename) {
    std::map<std::string, ByteImage::Ptr>::iterator it = refshim_registry().find(filename);
    if (it == refshim_registry().end()) throw util::FileException(filename, "not registered");
    ImageHeaders h = { it->second->width(), it->second->height(), it->second->channels(), IMAGE_TYPE_UINT8 };
    return h;
}
// "saving" registers the image under its file name so the glue can read it back (validity masks)
inline void save_png_file(ByteImage::ConstPtr img, std::string const& filename) { refshim_registry()[filename] = img->duplicate(); }
inline void save_file(ByteImage::ConstPtr img, std::string const& filename) { save_png_file(img, filename); }

} }  // namespace mve::image
