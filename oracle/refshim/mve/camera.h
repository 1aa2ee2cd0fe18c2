// refshim: MVE mve::CameraInfo stand-in (see ../README.md).  The glue fills the four quantities
// TextureView's constructor asks for directly; MVE's own derivation from focal length / rotation /
// translation is outside libs/tex and not restated.
#pragma once
#include <cstring>

namespace mve {

struct CameraInfo {
    float calibration[9];   // row major 3x3, already scaled to the image size
    float position[3];
    float viewdir[3];
    float world_to_cam[16]; // row major 4x4
    float flen;
    CameraInfo() : flen(1.0f) { std::memset(calibration, 0, sizeof(calibration)); std::memset(position, 0, sizeof(position));
                                std::memset(viewdir, 0, sizeof(viewdir)); std::memset(world_to_cam, 0, sizeof(world_to_cam)); }
    void fill_calibration(float* mat, float /*width*/, float /*height*/) const { std::memcpy(mat, calibration, sizeof(calibration)); }
    void fill_camera_pos(float* pos) const { std::memcpy(pos, position, sizeof(position)); }
    void fill_viewing_direction(float* dir) const { std::memcpy(dir, viewdir, sizeof(viewdir)); }
    void fill_world_to_cam(float* mat) const { std::memcpy(mat, world_to_cam, sizeof(world_to_cam)); }
};

}  // namespace mve
