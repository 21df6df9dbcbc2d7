// pcl_harness.cpp — the reference's own PCL call sequence, ROS-free (TEST INFRASTRUCTURE; built only where a PCL install exists).
//
// PCL / Eigen / FLANN / Boost are not available offline, so the oracle (oracle/cuboid_oracle.cpp) is "parity
// unpinned". This file is what pins it the day a PCL install is found (oracle/pcl_probe.py looks for one at bench time): it
// executes, literally and in order, the calls of
//     cuboid_detection/src/ground_plane_segmentation.cpp:53-101   (PassThrough z, PassThrough x, VoxelGrid, SACSegmentation, ExtractIndices)
//     cuboid_detection/src/iterative_closest_point.cpp:159-182    (loadPCDFile, IterativeClosestPoint::align, getFitnessScore)
// on a cloud read from a file, dumps every intermediate result the oracle also produces, and times the callback bodies
// (3 warm-ups, median of >= 10 runs: BASELINE.md section 3). Nothing here is copied from the reference: the calls are the
// public PCL API with the reference's constants.
//
// build:  g++ -O2 -std=c++14 pcl_harness.cpp $(pkg-config --cflags --libs pcl_registration pcl_segmentation pcl_filters pcl_io pcl_common) -o pcl_harness
// usage:  pcl_harness <cloud.bin> <template.pcd> <out.bin> [voxel_size distance_threshold icp_fitness_score runs]
//   cloud.bin : int32 n, then n * (x, y, z) float32  (the unprojected, unfiltered depth cloud)
//   out.bin   : int32 n_pass, n_vox; n_vox * xyz; int32 found; 4 float coefficients; int32 n_inl; n_inl * int32;
//               int32 n_remain; n_remain * xyz; 16 float final transform (row-major); double fitness; int32 converged;
//               double median_ms_segmentation; double median_ms_icp
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <vector>

#include <pcl/PCLPointCloud2.h>
#include <pcl/ModelCoefficients.h>
#include <pcl/PointIndices.h>
#include <pcl/conversions.h>
#include <pcl/filters/extract_indices.h>
#include <pcl/filters/passthrough.h>
#include <pcl/filters/voxel_grid.h>
#include <pcl/io/pcd_io.h>
#include <pcl/point_types.h>
#include <pcl/registration/icp.h>
#include <pcl/segmentation/sac_segmentation.h>

struct SegOut {
    pcl::PCLPointCloud2::Ptr voxel, remain;
    pcl::ModelCoefficients::Ptr coefficients;
    pcl::PointIndices::Ptr inliers;
    int n_pass = 0;
};

// ground_plane_segmentation.cpp:53-101
static SegOut segmentation(const pcl::PCLPointCloud2::Ptr& cloud_ptr, float voxel_size, double distance_threshold, bool invert) {
    SegOut o;
    pcl::PCLPointCloud2::Ptr cloud_filtered_ptr_z(new pcl::PCLPointCloud2), cloud_filtered_ptr(new pcl::PCLPointCloud2);
    pcl::PassThrough<pcl::PCLPointCloud2> pass_z;
    pass_z.setInputCloud(cloud_ptr);
    pass_z.setFilterFieldName("z");
    pass_z.setFilterLimits(0.0, 0.9);
    pass_z.filter(*cloud_filtered_ptr_z);
    pcl::PassThrough<pcl::PCLPointCloud2> pass;
    pass.setInputCloud(cloud_filtered_ptr_z);
    pass.setFilterFieldName("x");
    pass.setFilterLimits(-0.2, 0.2);
    pass.filter(*cloud_filtered_ptr);
    o.n_pass = (int)(cloud_filtered_ptr->width * cloud_filtered_ptr->height);
    o.voxel.reset(new pcl::PCLPointCloud2);
    pcl::VoxelGrid<pcl::PCLPointCloud2> downsample;
    downsample.setInputCloud(cloud_filtered_ptr);
    downsample.setLeafSize(voxel_size, voxel_size, voxel_size);
    downsample.filter(*o.voxel);
    o.coefficients.reset(new pcl::ModelCoefficients);
    o.inliers.reset(new pcl::PointIndices);
    pcl::SACSegmentation<pcl::PointXYZ> seg;
    pcl::PointCloud<pcl::PointXYZ>::Ptr pcl_voxel_ptr(new pcl::PointCloud<pcl::PointXYZ>);
    pcl::fromPCLPointCloud2(*o.voxel, *pcl_voxel_ptr);
    seg.setOptimizeCoefficients(true);
    seg.setModelType(pcl::SACMODEL_PLANE);
    seg.setMethodType(pcl::SAC_RANSAC);
    seg.setMaxIterations(1000);
    seg.setDistanceThreshold(distance_threshold);
    seg.setInputCloud(pcl_voxel_ptr);
    seg.segment(*o.inliers, *o.coefficients);
    o.remain.reset(new pcl::PCLPointCloud2);
    pcl::ExtractIndices<pcl::PCLPointCloud2> extract;
    extract.setInputCloud(o.voxel);
    extract.setIndices(o.inliers);
    extract.setNegative(invert);
    extract.filter(*o.remain);
    return o;
}

struct IcpOut { Eigen::Matrix4f T = Eigen::Matrix4f::Identity(); double fitness = 0.0; bool converged = false; };

// iterative_closest_point.cpp:170-182
static IcpOut registration(const pcl::PointCloud<pcl::PointXYZ>::Ptr& input_cuboid, const pcl::PointCloud<pcl::PointXYZ>::Ptr& template_cuboid,
                           double icp_fitness_score) {
    IcpOut o;
    pcl::PointCloud<pcl::PointXYZ> output_cloud;
    pcl::IterativeClosestPoint<pcl::PointXYZ, pcl::PointXYZ> icp;
    icp.setInputSource(input_cuboid);
    icp.setInputTarget(template_cuboid);
    icp.setMaximumIterations(5000);
    icp.setTransformationEpsilon(1e-9);
    icp.setEuclideanFitnessEpsilon(icp_fitness_score);
    icp.setRANSACOutlierRejectionThreshold(1.5);
    icp.align(output_cloud);
    o.T = icp.getFinalTransformation();
    o.converged = icp.hasConverged();
    o.fitness = icp.getFitnessScore();
    return o;
}

template <typename F>
static double median_ms(F&& f, int warm, int runs) {
    std::vector<double> t;
    for (int i = 0; i < warm + runs; ++i) {
        const auto a = std::chrono::steady_clock::now();
        f();
        const auto b = std::chrono::steady_clock::now();
        if (i >= warm) t.push_back(std::chrono::duration<double, std::milli>(b - a).count());
    }
    std::sort(t.begin(), t.end());
    return t.empty() ? 0.0 : t[t.size() / 2];
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: pcl_harness cloud.bin template.pcd out.bin [voxel thr fitness runs]\n"); return 2; }
    const float voxel_size = argc > 4 ? (float)atof(argv[4]) : 0.005f;
    const double distance_threshold = argc > 5 ? atof(argv[5]) : 0.015;
    const double icp_fitness_score = argc > 6 ? atof(argv[6]) : 0.0004;
    const int runs = argc > 7 ? atoi(argv[7]) : 10;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 3;
    int32_t n = 0;
    if (fread(&n, 4, 1, f) != 1) return 3;
    pcl::PointCloud<pcl::PointXYZ> in;
    in.resize((size_t)n);
    for (int i = 0; i < n; ++i) {
        float v[3];
        if (fread(v, 4, 3, f) != 3) return 3;
        in.points[(size_t)i].x = v[0]; in.points[(size_t)i].y = v[1]; in.points[(size_t)i].z = v[2];
    }
    fclose(f);
    pcl::PCLPointCloud2::Ptr cloud_ptr(new pcl::PCLPointCloud2);
    pcl::toPCLPointCloud2(in, *cloud_ptr);
    pcl::PointCloud<pcl::PointXYZ>::Ptr template_cuboid(new pcl::PointCloud<pcl::PointXYZ>);
    if (pcl::io::loadPCDFile<pcl::PointXYZ>(argv[2], *template_cuboid) == -1) { fprintf(stderr, "Couldn't read the template PCL file\n"); return 4; }

    SegOut s = segmentation(cloud_ptr, voxel_size, distance_threshold, true);
    pcl::PointCloud<pcl::PointXYZ>::Ptr input_cuboid(new pcl::PointCloud<pcl::PointXYZ>), vox(new pcl::PointCloud<pcl::PointXYZ>);
    pcl::fromPCLPointCloud2(*s.remain, *input_cuboid);
    pcl::fromPCLPointCloud2(*s.voxel, *vox);
    IcpOut r = registration(input_cuboid, template_cuboid, icp_fitness_score);
    const double ms_seg = median_ms([&] { segmentation(cloud_ptr, voxel_size, distance_threshold, true); }, 3, runs);
    const double ms_icp = median_ms([&] { registration(input_cuboid, template_cuboid, icp_fitness_score); }, 3, runs);

    FILE* o = fopen(argv[3], "wb");
    if (!o) return 5;
    auto put_i = [&](int32_t v) { fwrite(&v, 4, 1, o); };
    put_i(s.n_pass); put_i((int32_t)vox->size());
    for (const auto& p : vox->points) { const float v[3] = {p.x, p.y, p.z}; fwrite(v, 4, 3, o); }
    put_i(s.coefficients->values.size() == 4 ? 1 : 0);
    float c[4] = {0, 0, 0, 0};
    for (size_t k = 0; k < s.coefficients->values.size() && k < 4; ++k) c[k] = s.coefficients->values[k];
    fwrite(c, 4, 4, o);
    put_i((int32_t)s.inliers->indices.size());
    for (int idx : s.inliers->indices) put_i(idx);
    put_i((int32_t)input_cuboid->size());
    for (const auto& p : input_cuboid->points) { const float v[3] = {p.x, p.y, p.z}; fwrite(v, 4, 3, o); }
    float T[16];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T[4 * i + j] = r.T(i, j);
    fwrite(T, 4, 16, o);
    fwrite(&r.fitness, 8, 1, o);
    put_i(r.converged ? 1 : 0);
    fwrite(&ms_seg, 8, 1, o); fwrite(&ms_icp, 8, 1, o);
    fclose(o);
    printf("{\"n_pass\": %d, \"n_vox\": %zu, \"n_inliers\": %zu, \"n_remain\": %zu, \"converged\": %d, \"fitness\": %.9g, \"ms_segmentation\": %.3f, \"ms_icp\": %.3f}\n",
           s.n_pass, vox->size(), s.inliers->indices.size(), input_cuboid->size(), (int)r.converged, r.fitness, ms_seg, ms_icp);
    return 0;
}
