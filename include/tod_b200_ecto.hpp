// tod_b200_ecto.hpp — the two ecto cells of TOD's detection pipeline, re-implemented on libtod_b200.so.
//
// Drop-in for src/detection/DescriptorMatcher.cpp and src/detection/GuessGenerator.cpp of wg-perception/tod 0.5.6:
// same cell names ("DescriptorMatcher", "GuessGenerator") in the same boost.python module (ecto_detection,
// src/detection/module.cpp:38), same parameters, inputs and outputs, so python/object_recognition_tod/detector.py and
// conf/detection.ork work unchanged.  Compiled only when the ecto / ORK / OpenCV headers are there:
//
//     // src/detection/cells_b200.cpp  (replaces DescriptorMatcher.cpp and GuessGenerator.cpp in src/CMakeLists.txt:4-12)
//     #define TOD_WITH_ECTO 1
//     #include <tod_b200_ecto.hpp>
//     // target_link_libraries(ecto_detection_ectomodule tod_b200)
//
// Nothing here computes: the cells copy cv:: containers into the flat arrays of the C-ABI (cv::DMatch == tod_match and
// cv::KeyPoint == tod_keypoint field for field) and copy the results back.  This image has neither ecto nor ORK nor
// OpenCV C++ headers, so the guarded part is not built by the repo's own build; the unguarded layer underneath it
// (tod_b200.hpp) is, and is exercised from C++ by tests/cpp/cells_smoke.cpp.
#ifndef TOD_B200_ECTO_HPP_
#define TOD_B200_ECTO_HPP_

#include "tod_b200.hpp"

#ifdef TOD_WITH_ECTO

#include <ecto/ecto.hpp>
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>

#include <object_recognition_core/common/pose_result.h>
#include <object_recognition_core/common/types.h>
#include <object_recognition_core/db/ModelReader.h>
#include <object_recognition_core/db/opencv.h>

namespace tod
{
  using object_recognition_core::common::PoseResult;
  using object_recognition_core::db::ObjectId;

  static_assert(sizeof(cv::DMatch) == sizeof(tod_match), "cv::DMatch and tod_match must agree field for field");
  static_assert(sizeof(cv::KeyPoint) == sizeof(tod_keypoint), "cv::KeyPoint and tod_keypoint must agree");

  /** DescriptorMatcher (reference: src/detection/DescriptorMatcher.cpp:58-270) */
  struct DescriptorMatcher: public object_recognition_core::db::bases::ModelReaderBase
  {
    /** :60-129 — one DB document of method "TOD" per object; the span (:106-121) is computed by the library */
    void
    parameter_callback(const Documents & db_documents)
    {
      descriptors_db_.clear();
      features3d_db_.clear();
      std::vector<tod_b200::Document> docs;
      std::vector<std::string> ids;
      for (Documents::const_iterator document = db_documents.begin(); document != db_documents.end(); ++document)
      {
        ids.push_back(document->get_field<std::string>("object_id"));        // :72
        cv::Mat descriptors, points3d;
        document->get_attachment<cv::Mat>("descriptors", descriptors);        // :74-76
        document->get_attachment<cv::Mat>("points", points3d);                // :82
        if (points3d.rows != 1)
          points3d = points3d.t();                                            // :84-85
        descriptors_db_.push_back(descriptors.isContinuous() ? descriptors : descriptors.clone());
        features3d_db_.push_back(points3d.isContinuous() ? points3d : points3d.clone());
      }
      for (size_t i = 0; i < ids.size(); ++i)
      {
        tod_b200::Document d;
        d.object_id = ids[i];
        d.descriptors = descriptors_db_[i].ptr<uint8_t>();
        d.points = features3d_db_[i].ptr<float>();
        d.n = descriptors_db_[i].rows;
        docs.push_back(d);
      }
      impl_.parameter_callback(docs);                                         // clear + add + train (:127-128)
    }

    static void
    declare_params(ecto::tendrils& p)
    {
      object_recognition_core::db::bases::declare_params_impl(p, "TOD");      // :134
      p.declare < std::string > ("search_json_params",
                                 "JSON string that can contain the following fields: \"radius\" (for epsilon nearest "
                                 "neighbor search), \"ratio\" when applying the ratio criterion like in SIFT").required(true);
      p.declare<int>("device", "CUDA device ordinal (tod_b200 extension)", 0);
    }

    static void
    declare_io(const ecto::tendrils& params, ecto::tendrils& inputs, ecto::tendrils& outputs)
    {
      inputs.declare < cv::Mat > ("descriptors", "The descriptors to match to the database");
      outputs.declare < std::vector<std::vector<cv::DMatch> > > ("matches", "The matches for the input descriptors");
      outputs.declare < std::vector<cv::Mat> > ("matches_3d",
          "For each point, the 3d position of the matches, 1 by n matrix with 3 channels for, x, y, and z.");
      outputs.declare < std::vector<ObjectId> > ("object_ids", "The ids of the objects");
      outputs.declare < std::map<ObjectId, float> > ("spans", "The ids of the objects");
    }

    void
    configure(const ecto::tendrils& params, const ecto::tendrils& inputs, const ecto::tendrils& outputs)
    {
      // the matcher must exist before the DB callback fires (the reference creates it after configure_impl and relies
      // on the callback being deferred; here the order is made explicit)
      impl_.configure(params.get < std::string > ("search_json_params"), params.get<int>("device"));   // :159-181
      configure_impl();                                                                                // :157
    }

    int
    process(const ecto::tendrils& inputs, const ecto::tendrils& outputs)
    {
      const cv::Mat & descriptors_in = inputs.get < cv::Mat > ("descriptors");
      const cv::Mat descriptors = descriptors_in.isContinuous() ? descriptors_in : descriptors_in.clone();
      if (impl_.object_ids().empty())
      {
        std::cerr << "No descriptors loaded" << std::endl;                    // :204-208
        return ecto::OK;
      }
      const tod_b200::DescriptorMatcher::Outputs & out = impl_.process(descriptors.ptr<uint8_t>(), descriptors.rows);
      std::vector < std::vector<cv::DMatch> > matches(out.matches.size());
      std::vector < cv::Mat > matches_3d(out.matches.size());
      for (size_t q = 0; q < out.matches.size(); ++q)
      {
        const std::vector<tod_match> & src = out.matches[q];
        matches[q].resize(src.size());
        for (size_t j = 0; j < src.size(); ++j)
          matches[q][j] = cv::DMatch(src[j].queryIdx, src[j].trainIdx, src[j].imgIdx, src[j].distance);
        matches_3d[q] = cv::Mat(1, int(src.size()), CV_32FC3);               // :236
        if (!src.empty())
          std::memcpy(matches_3d[q].ptr<float>(), out.matches_3d[q].data(), src.size() * 3 * sizeof(float));
      }
      outputs["matches"] << matches;                                          // :246-249
      outputs["matches_3d"] << matches_3d;
      outputs["object_ids"] << out.object_ids;
      outputs["spans"] << out.spans;
      return ecto::OK;
    }

    tod_b200::DescriptorMatcher impl_;
    std::vector<cv::Mat> descriptors_db_, features3d_db_;   // keep the attachments alive while they are being copied
  };

  /** GuessGenerator (reference: src/detection/GuessGenerator.cpp:69-276) */
  struct GuessGenerator
  {
    static void
    declare_params(ecto::tendrils& params)
    {
      params.declare(&GuessGenerator::min_inliers_, "min_inliers", "Minimum number of inliers", 15);
      params.declare(&GuessGenerator::n_ransac_iterations_, "n_ransac_iterations", "Number of RANSAC iterations.", 1000);
      params.declare(&GuessGenerator::sensor_error_, "sensor_error", "The error (in meters) from the Kinect", 0.01);
      params.declare(&GuessGenerator::visualize_, "visualize", "If true, display temporary info through highgui", false);
      params.declare(&GuessGenerator::json_db_, "db", "The DB to get data from, as a JSON string").required(true);
      params.declare<int>("device", "CUDA device ordinal (tod_b200 extension)", 0);
      params.declare<int>("seed", "Seed of the RANSAC sampler stream (tod_b200 extension; the reference uses the "
                          "unseeded libc rand())", 0);
    }

    static void
    declare_io(const ecto::tendrils& params, ecto::tendrils& inputs, ecto::tendrils& outputs)
    {
      inputs.declare<cv::Mat>("image", "The height by width 3 channel point cloud");
      inputs.declare<cv::Mat>("points3d", "The height by width 3 channel point cloud");
      inputs.declare<std::vector<cv::KeyPoint> >("keypoints", "The interesting keypoints");
      inputs.declare<std::vector<std::vector<cv::DMatch> > >("matches", "The list of OpenCV DMatch");
      inputs.declare<std::vector<cv::Mat> >("matches_3d",
          "The corresponding 3d position of those matches. For each point, a 1 by n 3 channel matrix (for x,y and z)");
      inputs.declare<std::map<ObjectId, float> >("spans", "For each found object, its span based on known features.");
      inputs.declare<std::vector<ObjectId> >("object_ids", "The ids used in the matches");
      outputs.declare(&GuessGenerator::pose_results_, "pose_results", "The results of object recognition");
      outputs.declare(&GuessGenerator::Rs_, "Rs", "The rotations of the poses (useful for visualization)");
      outputs.declare(&GuessGenerator::Ts_, "Ts", "The translations of the poses (useful for visualization)");
    }

    void
    configure(const ecto::tendrils& params, const ecto::tendrils& inputs, const ecto::tendrils& outputs)
    {
      // `visualize` is accepted and ignored: the highgui debug drawing (:160-166, :210-221, :237-241) is not part of
      // the hot path.  The DB handle is only stamped into the PoseResults (:228).
      db_ = object_recognition_core::db::ObjectDbParameters(*json_db_).generateDb();                  // :119
      impl_.configure(*min_inliers_, *n_ransac_iterations_, *sensor_error_, params.get<int>("device"),
                      uint64_t(params.get<int>("seed")));
    }

    int
    process(const ecto::tendrils& inputs, const ecto::tendrils& outputs)
    {
      const std::vector<std::vector<cv::DMatch> > & matches = inputs.get<std::vector<std::vector<cv::DMatch> > >("matches");
      const std::vector<cv::Mat> & matches_3d = inputs.get<std::vector<cv::Mat> >("matches_3d");
      const std::vector<cv::KeyPoint> & keypoints = inputs.get<std::vector<cv::KeyPoint> >("keypoints");
      const cv::Mat point_cloud_in = inputs.get<cv::Mat>("points3d");
      const std::vector<ObjectId> & object_ids_in = inputs.get<std::vector<ObjectId> >("object_ids");
      const std::map<ObjectId, float> & spans = inputs.get<std::map<ObjectId, float> >("spans");
      pose_results_->clear();                                                 // :144-146
      Rs_->clear();
      Ts_->clear();
      if (point_cloud_in.empty())
        return ecto::OK;                                                      // :147-152 (2d-3d matching is a TODO there)
      const cv::Mat point_cloud = point_cloud_in.isContinuous() ? point_cloud_in : point_cloud_in.clone();

      // flat views of the inputs: row stride k = the longest match list
      size_t k = 1;
      for (size_t q = 0; q < matches.size(); ++q)
        k = std::max(k, matches[q].size());
      std::vector<tod_match> flat(matches.size() * k);
      std::vector<int32_t> counts(matches.size());
      std::vector<float> pts(matches.size() * k * 3, 0.f);
      for (size_t q = 0; q < matches.size(); ++q)
      {
        counts[q] = int32_t(matches[q].size());
        for (size_t j = 0; j < matches[q].size(); ++j)
        {
          const cv::DMatch & m = matches[q][j];
          tod_match t = { m.queryIdx, m.trainIdx, m.imgIdx, m.distance };
          flat[q * k + j] = t;
          const cv::Vec3f & p = matches_3d[q].at<cv::Vec3f>(0, int(j));
          pts[(q * k + j) * 3] = p[0];
          pts[(q * k + j) * 3 + 1] = p[1];
          pts[(q * k + j) * 3 + 2] = p[2];
        }
      }
      std::vector<float> spans_by_index(object_ids_in.size(), 0.f);
      for (size_t o = 0; o < object_ids_in.size(); ++o)
        spans_by_index[o] = spans.find(object_ids_in[o])->second;            // :189
      std::vector<tod_pose> poses(64 * std::max<size_t>(1, object_ids_in.size()));
      int32_t n_poses = 0;
      tod_b200::check(tod_guess_process(impl_.handle(), reinterpret_cast<const tod_keypoint *>(keypoints.data()),
                                        int32_t(keypoints.size()), point_cloud.ptr<float>(), point_cloud.rows,
                                        point_cloud.cols, flat.data(), counts.data(), int32_t(k), pts.data(),
                                        spans_by_index.data(), int32_t(spans_by_index.size()), poses.data(),
                                        int32_t(poses.size()), &n_poses, NULL, 0));
      for (int32_t i = 0; i < n_poses; ++i)
      {
        cv::Matx33f R_mat(poses[i].R);
        cv::Vec3f tvec(poses[i].T[0], poses[i].T[1], poses[i].T[2]);
        PoseResult pose_result;                                               // :224-230
        pose_result.set_R(cv::Mat(R_mat));
        pose_result.set_T(cv::Mat(tvec));
        pose_result.set_object_id(db_, object_ids_in[size_t(poses[i].object_index)]);
        pose_results_->push_back(pose_result);
        Rs_->push_back(cv::Mat(R_mat));
        Ts_->push_back(cv::Mat(tvec));
      }
      return ecto::OK;
    }

  private:
    ecto::spore<unsigned int> min_inliers_, n_ransac_iterations_;            // :255-269
    ecto::spore<float> sensor_error_;
    ecto::spore<bool> visualize_;
    ecto::spore<std::string> json_db_;
    ecto::spore<std::vector<PoseResult> > pose_results_;
    ecto::spore<std::vector<cv::Mat> > Rs_, Ts_;
    object_recognition_core::db::ObjectDbPtr db_;
    tod_b200::GuessGenerator impl_;
  };
}

ECTO_CELL(ecto_detection, tod::DescriptorMatcher, "DescriptorMatcher",
          "Given descriptors, find matches, relating to objects (exact Hamming k-NN on a B200).")       // :269
ECTO_CELL(ecto_detection, tod::GuessGenerator, "GuessGenerator",
          "Given matches and 3d positions, compute object poses (adjacency + RANSAC on a B200).")       // :275

#endif  // TOD_WITH_ECTO
#endif  // TOD_B200_ECTO_HPP_
