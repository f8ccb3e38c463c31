// Depth-filter object shared between depth.cu and tracker.cu (the tracker feeds it and is fed by it).
#pragma once
#include "common.cuh"

struct edsgpu_depth_points {
    edsgpu_ctx* ctx = nullptr;
    int N = 0;
    double fx = 0, fy = 0, cx = 0, cy = 0, mu_range = 0, px_error_angle = 0;
    double* state = nullptr;   // [N][4]
    double* coords = nullptr;  // [2][N][2] staging: kf, ef
    unsigned char* ok = nullptr;        // filter result per point: 1 = updated and valid
    unsigned char* outlier = nullptr;   // Tracker::getCoord: 1 = the point projects outside the event frame (its own buffer: the
                                        // flags have the opposite sense of `ok`)
    bool kf_coord_set = false;  // coords[0..2N) holds KeyFrame::coord
};

// DepthPoints::update with the pose taken from a tracker's 14-double state on the device and the event-frame
// coordinates already on the device (Tracker::getCoord); asynchronous on the context's stream.
edsgpu_status edsgpu_depth_update_tracked(edsgpu_depth_points* dp, const double* tracker_state_dev, const double* kf_coord_dev,
                                          const double* ef_coord_dev);
