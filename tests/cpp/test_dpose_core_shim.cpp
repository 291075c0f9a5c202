// test_dpose_core_shim.cpp — the reference's gtests restated against include/dpose_core/*.hpp.
//
// No gtest in this image: a tiny CHECK macro counts failures and the process exits non-zero on any.
//   ./test_dpose_core_shim --cpu-only   host-side contracts only (no CUDA device needed)
//   ./test_dpose_core_shim              everything (needs a GPU)
// Each block names the upstream test it restates (paths relative to dpose_core/test/).
#include <dpose_core/dpose_costmap.hpp>
#include <dpose_core/dpose_solve.hpp>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <random>
#include <set>

using namespace dpose_core;

static int g_checks = 0, g_failed = 0;
#define CHECK(cond)                                                       \
  do {                                                                    \
    ++g_checks;                                                           \
    if (!(cond)) {                                                        \
      ++g_failed;                                                         \
      std::fprintf(stderr, "%s:%d: CHECK failed: %s\n", __FILE__, __LINE__, #cond); \
    }                                                                     \
  } while (0)

static polygon
make_polygon(std::initializer_list<std::pair<int, int>> pts) {
  polygon p(2, (int)pts.size());
  int i = 0;
  for (const auto& pt : pts) {
    p(0, i) = pt.first;
    p(1, i) = pt.second;
    ++i;
  }
  return p;
}

static polygon
make_ship() {  // jacobian_data.cpp:8-22
  return make_polygon({{50, 0}, {40, -20}, {-30, -20}, {-60, 0}, {-30, 20}, {40, 20}});
}

static polygon
make_triangle() {  // cost_data.cpp:10-20
  return make_polygon({{-2, -1}, {2, -1}, {0, 3}});
}

static polygon
shifted(const polygon& p, int dx, int dy) {
  polygon q = p;
  for (int i = 0; i < q.cols(); ++i) {
    q(0, i) += dx;
    q(1, i) += dy;
  }
  return q;
}

template <typename F>
static bool
throws_runtime_error(F&& f, const char* needle) {
  try {
    f();
  } catch (const std::runtime_error& e) {
    return std::strstr(e.what(), needle) != nullptr;
  } catch (...) {
    return false;
  }
  return false;
}

// ---- host-side contracts ---------------------------------------------------------------------------------
static void
test_host() {
  // pose_gradient.cpp:9-17: fewer than three points throw
  for (int n = 0; n < 3; ++n) {
    polygon p(2, n);
    CHECK(throws_runtime_error([&] { pose_gradient pg(p, {}); }, "at least three points"));
  }
  // dpose_costmap.cpp:44-45: resolution <= 0 throws
  std::vector<geometry_msgs::Point> fp(4);
  fp[0].x = 0.5, fp[0].y = 0.3, fp[1].x = -0.5, fp[1].y = 0.3, fp[2].x = -0.5, fp[2].y = -0.3, fp[3].x = 0.5, fp[3].y = -0.3;
  CHECK(throws_runtime_error([&] { make_footprint(fp, 0.0); }, "resolution must be positive"));
  CHECK(throws_runtime_error([&] { make_footprint(fp, -0.05); }, "resolution must be positive"));
  const polygon cells = make_footprint(fp, 0.05);
  CHECK(cells.cols() == 4 && cells(0, 0) == 10 && cells(1, 0) == 6 && cells(0, 2) == -10 && cells(1, 2) == -6);
  // to_rectangle (dpose_core.hpp:45-69)
  const auto r = to_rectangle(3, 4);
  CHECK(r(0, 0) == 0 && r(0, 1) == 3 && r(1, 2) == 4 && r(0, 3) == 0 && r(0, 4) == 0 && r(1, 4) == 0);
  const auto r2 = to_rectangle(Eigen::Vector2i{1, 2}, Eigen::Vector2i{5, 7});
  CHECK(r2(0, 0) == 1 && r2(1, 0) == 2 && r2(0, 2) == 5 && r2(1, 2) == 7 && r2(0, 4) == 1);
  // raytrace: end exclusive, max(|dx|,|dy|) cells, starts at begin (dpose_costmap.cpp:57-104)
  const cell_vector ray = raytrace(cell(0, 0), cell(5, 2));
  CHECK(ray.size() == 5 && ray.front() == cell(0, 0));
  for (size_t i = 1; i < ray.size(); ++i) CHECK(ray[i].x() == ray[i - 1].x() + 1);
  CHECK(raytrace(cell(3, 3), cell(3, 3)).empty());
  // to_rays of an axis-aligned ring: every row of the box spans [0, w]
  const cell_rays rays = to_rays(to_rectangle(6, 3));
  CHECK(rays.size() == 4);
  for (const auto& kv : rays) CHECK(kv.second.min == 0 && kv.second.max == 6);
  // x_bresenham follows raytrace row by row (check_footprint.cpp:92-146) and starts at the lower end (:148-160)
  for (int k = 1; k < 14; ++k) {
    const double ang = 0.1 * k;
    const cell zero(0, 0), end((int)(std::cos(ang) * 100), (int)(std::sin(ang) * 100));
    x_bresenham br(zero, end);
    int curr_y = -1;
    for (const auto& pt : raytrace(zero, end))
      if (pt.y() != curr_y) {
        CHECK(pt.x() == br.get_x());
        br.advance_x();
        curr_y = pt.y();
      }
  }
  CHECK(x_bresenham(cell(0, 0), cell(10, 12)).get_x() == x_bresenham(cell(10, 12), cell(0, 0)).get_x());
  {  // is_inside (check_footprint.cpp:23-88)
    costmap_2d::Costmap2D cm(30, 40, 0.1, 0, 0);
    CHECK(is_inside(cm, polygon()));
    polygon fit = make_polygon({{0, 0}, {29, 0}, {29, 39}, {0, 39}});
    CHECK(is_inside(cm, fit));
    const int pert[8][2] = {{0, -1}, {1, -1}, {2, 1}, {3, -1}, {4, 1}, {5, 1}, {6, -1}, {7, 1}};
    for (const auto& pp : pert) {
      polygon moved = fit;
      moved(pp[0] % 2, pp[0] / 2) += pp[1];
      CHECK(!is_inside(cm, moved));
    }
  }
  // default-constructed objects are inert
  pose_gradient empty;
  CHECK(throws_runtime_error([&] { cell_vector c; empty.get_cost({0, 0, 0}, c.cbegin(), c.cend(), nullptr); }, "not initialised"));
}

// ---- cost_data geometry (cost_data.cpp) ------------------------------------------------------------------------
static void
test_cost_data() {
  const polygon tri = make_triangle();
  const cost_data original(tri, 0);
  const int shifts[4][2] = {{10, 0}, {4, 5}, {-4, -5}, {0, -9}};  // cost_data.cpp:23-62
  for (const auto& s : shifts) {
    const cost_data moved(shifted(tri, s[0], s[1]), 0);
    CHECK(moved.get_data().rows == original.get_data().rows && moved.get_data().cols == original.get_data().cols);
    for (int i = 0; i < 5; ++i) {
      CHECK(moved.get_box()(0, i) == original.get_box()(0, i) + s[0]);
      CHECK(moved.get_box()(1, i) == original.get_box()(1, i) + s[1]);
    }
    CHECK(moved.get_center().x() == original.get_center().x() - s[0]);
    CHECK(moved.get_center().y() == original.get_center().y() - s[1]);
  }
  for (int pad : {1, 5, 10}) {  // cost_data.cpp:68-115
    const cost_data padded(tri, (size_t)pad);
    CHECK(padded.get_data().rows == original.get_data().rows + 2 * pad);
    CHECK(padded.get_data().cols == original.get_data().cols + 2 * pad);
    CHECK(padded.get_center().x() == original.get_center().x() + pad);
    CHECK(padded.get_center().y() == original.get_center().y() + pad);
    CHECK(padded.get_box()(0, 0) == original.get_box()(0, 0) - pad && padded.get_box()(0, 1) == original.get_box()(0, 1) + pad);
    CHECK(padded.get_box()(1, 0) == original.get_box()(1, 0) - pad && padded.get_box()(1, 2) == original.get_box()(1, 2) + pad);
  }
  for (int k = 0; k < 29; ++k) {  // cost_data.cpp:121-154: size follows the rotated footprint's extent
    const double ang = 0.1 + 0.1 * k, c = std::cos(ang), s = std::sin(ang);
    polygon rot(2, 3);
    int lo[2] = {1 << 30, 1 << 30}, hi[2] = {-(1 << 30), -(1 << 30)};
    for (int i = 0; i < 3; ++i) {
      const double x = 10.0 * tri(0, i), y = 10.0 * tri(1, i);
      rot(0, i) = (int)std::round(c * x - s * y);
      rot(1, i) = (int)std::round(s * x + c * y);
      for (int d = 0; d < 2; ++d) {
        lo[d] = std::min(lo[d], rot(d, i));
        hi[d] = std::max(hi[d], rot(d, i));
      }
    }
    const int pad = 3;
    const cost_data cd(rot, pad);
    CHECK(cd.get_data().cols == hi[0] - lo[0] + 1 + 2 * pad);
    CHECK(cd.get_data().rows == hi[1] - lo[1] + 1 + 2 * pad);
    CHECK(cd.get_box()(0, 0) == lo[0] - pad && cd.get_box()(0, 1) == hi[0] + 1 + pad);
    CHECK(cd.get_box()(1, 0) == lo[1] - pad && cd.get_box()(1, 2) == hi[1] + 1 + pad);
  }
  // table values are non-negative and smallest far outside (dpose_core/README.md:12-13)
  const cv::Mat& t = original.get_data();
  float mn = 1e30f;
  for (int r = 0; r < t.rows; ++r)
    for (int c = 0; c < t.cols; ++c) mn = std::min(mn, t.at<float>(r, c));
  CHECK(mn >= 0.f);
}

// ---- Jacobian self-consistency (jacobian_data.cpp:26-78, plus theta) ----------------------------------------------
static void
test_jacobian() {
  const pose_gradient pg(make_ship(), {10});
  cell_vector cells;
  for (int xx = 0; xx != 20; ++xx)
    for (int yy = 0; yy != 20; ++yy) cells.emplace_back(xx, yy);
  for (int k = 1; k < 15; ++k) {
    const double th = 0.1 * k;
    pose_gradient::jacobian J;
    const double cost = pg.get_cost({0, 0, th}, cells.cbegin(), cells.cend(), &J);
    CHECK(cost > 0);
    for (int d = 0; d < 3; ++d) {
      double lo[3] = {0, 0, th}, hi[3] = {0, 0, th};
      lo[d] -= 1e-6, hi[d] += 1e-6;
      const double left = pg.get_cost({lo[0], lo[1], lo[2]}, cells.cbegin(), cells.cend(), nullptr);
      const double right = pg.get_cost({hi[0], hi[1], hi[2]}, cells.cbegin(), cells.cend(), nullptr);
      const double diff = (right - left) / 2e-6;
      const double err = std::abs((diff - J(d)) / (diff != 0 ? diff : 1.));
      CHECK(err <= 1e-3);
    }
  }
  // empty range: cost 0, J zeroed (dpose_core.cpp:224-225)
  pose_gradient::jacobian J(1, 2, 3);
  cell_vector none;
  CHECK(pg.get_cost({0, 0, 0}, none.cbegin(), none.cend(), &J) == 0.0);
  CHECK(J(0) == 0.0 && J(1) == 0.0 && J(2) == 0.0);
}

// ---- lethal_cells_within (lethal_cells_within.cpp) -------------------------------------------------------------------
static std::set<std::pair<int, int>>
convex_fill_cells_ros(const cell_rectangle& ring) {
  // costmap_2d::Costmap2D::convexFillCells restated from its published algorithm: polygon outline via
  // raytraceLine, then every column filled between its smallest and largest y
  std::map<int, std::pair<int, int>> cols;
  for (int i = 0; i < 4; ++i)
    for (const auto& c : raytrace(cell(ring(0, i), ring(1, i)), cell(ring(0, i + 1), ring(1, i + 1)))) {
      auto it = cols.find(c.x());
      if (it == cols.end())
        cols[c.x()] = {c.y(), c.y()};
      else
        it->second = {std::min(it->second.first, c.y()), std::max(it->second.second, c.y())};
    }
  std::set<std::pair<int, int>> out;
  for (const auto& kv : cols)
    for (int y = kv.second.first; y <= kv.second.second; ++y) out.insert({kv.first, y});
  return out;
}

static void
test_lethal() {
  using costmap_2d::Costmap2D;
  using costmap_2d::LETHAL_OBSTACLE;
  {  // :29-43 dot
    Costmap2D map(200, 200, 0.1, 0, 0);
    map.setCost(100, 100, LETHAL_OBSTACLE);
    const auto cells = lethal_cells_within(map, to_rectangle(Eigen::Vector2i{99, 99}, Eigen::Vector2i{101, 101}));
    CHECK(cells.size() == 1 && cells.front() == cell(100, 100));
  }
  {  // :45-57 sparse line; 253 / 255 are not lethal
    Costmap2D map(200, 200, 0.1, 0, 0);
    for (unsigned ii = 0; ii < map.getSizeInCellsX(); ii += 20) map.setCost(ii, 100, LETHAL_OBSTACLE);
    map.setCost(90, 100, costmap_2d::INSCRIBED_INFLATED_OBSTACLE);
    map.setCost(91, 100, costmap_2d::NO_INFORMATION);
    const auto cells = lethal_cells_within(map, to_rectangle(Eigen::Vector2i{80, 99}, Eigen::Vector2i{121, 101}));
    CHECK(cells.size() == 3);
  }
  {  // empty map (dpose_costmap.cpp:124-125)
    Costmap2D map(0, 0, 0.1, 0, 0);
    CHECK(lethal_cells_within(map, to_rectangle(5, 5)).empty());
  }
  {  // :61-115 regression against convexFillCells on a fully lethal map, 20 rotations
    Costmap2D map(200, 200, 0.1, 0, 0, LETHAL_OBSTACLE);
    const auto box = to_rectangle(Eigen::Vector2i{-20, -10}, Eigen::Vector2i{30, 20});
    for (int k = 0; k < 20; ++k) {
      const double ang = 0.1 * k, c = std::cos(ang), s = std::sin(ang);
      cell_rectangle rot;
      for (int i = 0; i < 5; ++i) {
        rot(0, i) = (int)std::round(100 + c * box(0, i) - s * box(1, i));
        rot(1, i) = (int)std::round(100 + s * box(0, i) + c * box(1, i));
      }
      const auto cells = lethal_cells_within(map, rot);
      CHECK(!cells.empty());
      std::set<std::pair<int, int>> ours;
      for (const auto& cc : cells) ours.insert({cc.x(), cc.y()});
      CHECK(ours.size() == cells.size());
      CHECK(ours == convex_fill_cells_ros(rot));
    }
  }
}

// ---- check_footprint (check_footprint.cpp:162-212): regression against the ROS outline / convex fill ------------------
static void
test_check_footprint() {
  const polygon square = make_polygon({{10, -10}, {10, 8}, {-12, 9}, {-5, -7}});  // :137-147
  for (int k = 0; k < 63; k += 3) {
    const double ang = 0.1 * k, c = std::cos(ang), s = std::sin(ang);
    polygon fp(2, 4);
    cell_rectangle ring;
    for (int i = 0; i < 4; ++i) {
      fp(0, i) = (int)std::round(50 + c * square(0, i) - s * square(1, i));
      fp(1, i) = (int)std::round(50 + s * square(0, i) + c * square(1, i));
      ring(0, i) = fp(0, i), ring(1, i) = fp(1, i);
    }
    ring(0, 4) = fp(0, 0), ring(1, 4) = fp(1, 0);
    const auto area = convex_fill_cells_ros(ring);
    costmap_2d::Costmap2D cm(100, 100, 0.1, 0, 0, costmap_2d::LETHAL_OBSTACLE);
    for (const auto& xy : area) cm.setCost(xy.first, xy.second, costmap_2d::FREE_SPACE);
    CHECK(check_footprint(cm, fp, costmap_2d::LETHAL_OBSTACLE));
    int i = 0;
    for (int e = 0; e < 4; ++e)
      for (const auto& oc : raytrace(cell(ring(0, e), ring(1, e)), cell(ring(0, e + 1), ring(1, e + 1))))
        if (i++ % 7 == 0) {
          cm.setCost(oc.x(), oc.y(), costmap_2d::LETHAL_OBSTACLE);
          CHECK(!check_footprint(cm, fp, costmap_2d::LETHAL_OBSTACLE));
          cm.setCost(oc.x(), oc.y(), costmap_2d::FREE_SPACE);
        }
  }
}

// ---- batched get_cost == lethal_cells_within + get_cost per pose --------------------------------------------------------
static void
test_batch() {
  const int side = 400;
  std::mt19937 rng(7);
  costmap_2d::Costmap2D map(side, side, 0.05, 0, 0);
  for (int y = 0; y < side; ++y)
    for (int x = 0; x < side; ++x)
      if (rng() % 10 == 0) map.setCost(x, y, costmap_2d::LETHAL_OBSTACLE);
  const polygon rect = make_polygon({{10, 6}, {-10, 6}, {-10, -6}, {10, -6}});
  const pose_gradient pg(rect, {2});
  const device_costmap dmap(map.getCharMap(), side, side, side);
  std::uniform_real_distribution<double> ux(32.0, side - 32.0), ut(-M_PI, M_PI);
  const int n = 500;
  std::vector<pose_gradient::pose> poses;
  for (int i = 0; i < n; ++i) poses.emplace_back(ux(rng), ux(rng), ut(rng));
  std::vector<double> costs;
  std::vector<pose_gradient::jacobian> Js;
  pg.get_cost(dmap, poses, costs, &Js);
  CHECK((int)costs.size() == n && (int)Js.size() == n);
  const auto& box = pg.get_data().get_box();
  for (int i = 0; i < n; ++i) {
    const double c = std::cos(poses[i](2)), s = std::sin(poses[i](2));
    cell_rectangle rot;
    for (int k = 0; k < 5; ++k) {
      rot(0, k) = (int)std::round(poses[i](0) + c * box(0, k) - s * box(1, k));
      rot(1, k) = (int)std::round(poses[i](1) + s * box(0, k) + c * box(1, k));
    }
    const cell_vector cells = lethal_cells_within(dmap, rot);
    pose_gradient::jacobian J;
    const double cost = pg.get_cost(poses[i], cells.cbegin(), cells.cend(), &J);
    CHECK(std::abs(cost - costs[i]) <= 1e-6 + 1e-5 * std::abs(cost));
    for (int d = 0; d < 3; ++d) CHECK(std::abs(J(d) - Js[i](d)) <= 1e-6 + 1e-5 * std::abs(J(d)));
  }
}

// ---- the multi-start solve: preProcess's attempt loop (dpose_goal_tolerance.cpp:430-456) in one call ---------------------
static void
test_solve() {
  const int side = 300;
  costmap_2d::Costmap2D map(side, side, 0.05, 0, 0);
  for (int y = 145; y < 156; ++y)
    for (int x = 145; x < 156; ++x) map.setCost(x, y, costmap_2d::LETHAL_OBSTACLE);  // a block on the goal
  const polygon rect = make_polygon({{10, 6}, {-10, 6}, {-10, -6}, {10, -6}});
  const pose_gradient pg(rect, {2});
  const device_costmap dmap(map.getCharMap(), side, side, side);
  multi_start_problem pr(pg, dmap);
  const dpose_gt_config cfg = {20.0, 1.0, 1.0, 1.0, 1.0, 1.0};
  pr.init(pose_gradient::pose(100, 100, 0), pose_gradient::pose(150, 150, 0), cfg);
  const auto res = pr.solve(256, 9);
  CHECK(res.x.size() == 3 * 256 && res.f.size() == 256);
  CHECK(res.accepted() && res.summary.n_accepted > 0);
  {  // attempt 0 starts at the goal, where the footprint collides with the block
    polygon at_goal(2, 4);
    for (int k = 0; k < 4; ++k) at_goal(0, k) = 150 + rect(0, k), at_goal(1, k) = 150 + rect(1, k);
    CHECK(!check_footprint(map, at_goal, costmap_2d::LETHAL_OBSTACLE));
  }
  for (size_t i = 0; i < 256; ++i) {
    // every accepted start is one the upstream validation accepts (:386-399): transform, round, check_footprint
    const double px = 150 + res.x[3 * i], py = 150 + res.x[3 * i + 1], pt = res.x[3 * i + 2];
    polygon moved(2, 4);
    for (int k = 0; k < 4; ++k) {
      moved(0, k) = (int)std::round(px + std::cos(pt) * rect(0, k) - std::sin(pt) * rect(1, k));
      moved(1, k) = (int)std::round(py + std::sin(pt) * rect(0, k) + std::cos(pt) * rect(1, k));
    }
    const bool free_ = check_footprint(map, moved, costmap_2d::LETHAL_OBSTACLE);
    CHECK(free_ == ((res.status[i] & DPOSE_SOLVE_FOOTPRINT_FREE) != 0));
    CHECK(res.x[3 * i] * res.x[3 * i] + res.x[3 * i + 1] * res.x[3 * i + 1] <= 400.0 * (1 + 1e-12));
    CHECK(std::abs(res.x[3 * i + 2]) <= 1.0);
  }
}

int
main(int argc, char** argv) {
  const bool cpu_only = argc > 1 && !std::strcmp(argv[1], "--cpu-only");
  test_host();
  if (!cpu_only) {
    int ndev = 0;
    if (dpose_device_count(&ndev) != DPOSE_OK || ndev == 0) {
      std::fprintf(stderr, "no CUDA device: %s\n", dpose_last_error());
      return 2;
    }
    test_cost_data();
    test_jacobian();
    test_lethal();
    test_check_footprint();
    test_batch();
    test_solve();
  } else {
    // without a device the constructors must fail loudly, not fall back to a CPU path
    int ndev = 0;
    dpose_device_count(&ndev);
    if (ndev == 0) CHECK(throws_runtime_error([&] { cost_data cd(make_triangle(), 0); }, "no CPU fallback"));
  }
  std::printf("%d checks, %d failed\n", g_checks, g_failed);
  return g_failed ? 1 : 0;
}
