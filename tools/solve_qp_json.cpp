/*
 * solve_qp_json -- the wire format of the reference's benchmark harness, in front of
 * libtreeqp_b200.so.
 *
 * Same contract as the reference's examples/solve_qp_json.cpp:206-615 (what benchmark/utils/
 * treeqp_solve.m calls): read a tree QP as JSON
 *     nodes[] {Q, R, S, q, r, lx, ux, lu, uu}   edges[] {A, B, b, from, to}
 *     options {solver: "tdunes" | "sdunes", maxit, stationarityTolerance, lineSearchMaxIter,
 *              lineSearchBeta, lineSearchGamma, checkLastActiveSet, clipping, regType, regTol, regValue}
 *     init    {x0, lam0_tree | lam0_scen + mu0_scen}        (optionally from a second file)
 * eliminate x0, solve NREP times from the same dual start (checking that status and iteration
 * count repeat, cpp:428-429), and write
 *     solution {nodes[] {x, u, mu_x, mu_u}, edges[] {lam}}, info {solver, cpu_time, status, num_iter,
 *     kkt_tol, ls_iters}, init {...}
 * Matrices are arrays of rows; 1 x n and n x 1 matrices are flat arrays, 1 x 1 a number (what
 * MATLAB's jsonencode produces; readColMajorMatrix, cpp:92-108).  Missing bounds mean +-TREEQP_INF
 * (the examples/random_qp_utils/data0N.json files carry none).
 *
 *     solve_qp_json problem.json [init.json] [--out result.json] [--nrep N] [--dry-run]
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include <nlohmann/json.hpp>

#include "treeqp_b200.h"

using nlohmann::json;
typedef std::vector<double> dvec;

static dvec flat(const json &j, size_t n)
{   /* n numbers out of a number, a flat array or (n x 1 / 1 x n) nested arrays */
    dvec v;
    if (j.is_number()) v.push_back(j.get<double>());
    else
        for (const auto &e : j)
        {
            if (e.is_array()) for (const auto &f : e) v.push_back(f.get<double>());
            else v.push_back(e.get<double>());
        }
    if (v.size() != n) throw std::invalid_argument("vector of length " + std::to_string(v.size()) + ", expected " + std::to_string(n));
    return v;
}

static dvec colmajor(const json &j, size_t m, size_t n)
{
    if (m == 0 || n == 0) return dvec();
    if (m == 1 || n == 1) return flat(j, m * n);
    dvec v(m * n);
    for (size_t r = 0; r < m; r++)
        for (size_t c = 0; c < n; c++) v[r + c * m] = j.at(r).at(c).get<double>();
    return v;
}

static size_t count_of(const json &j) { return j.is_array() ? j.size() : (j.is_null() ? 0 : 1); }

static dvec bound(const json &node, const char *key, size_t n, double fill)
{
    if (!node.count(key) || node.at(key).is_null()) return dvec(n, fill);
    dvec v = flat(node.at(key), n);
    return v;
}

static int reg_type(const json &o)
{
    if (!o.count("regType")) return TREEQP_ON_THE_FLY_LEVENBERG_MARQUARDT;
    if (o["regType"].is_number()) return o["regType"].get<int>();
    const std::string s = o["regType"];
    if (s == "TREEQP_NO_REGULARIZATION") return TREEQP_NO_REGULARIZATION;
    if (s == "TREEQP_ALWAYS_LEVENBERG_MARQUARDT") return TREEQP_ALWAYS_LEVENBERG_MARQUARDT;
    if (s == "TREEQP_ON_THE_FLY_LEVENBERG_MARQUARDT") return TREEQP_ON_THE_FLY_LEVENBERG_MARQUARDT;
    throw std::invalid_argument("unknown regType " + s);
}

template <typename Opts>
static void common_options(const json &o, Opts &opts)
{
    if (o.count("maxit")) opts.maxIter = o["maxit"];
    if (o.count("stationarityTolerance")) opts.stationarityTolerance = o["stationarityTolerance"];
    if (o.count("lineSearchMaxIter")) opts.lineSearchMaxIter = o["lineSearchMaxIter"];
    if (o.count("lineSearchBeta")) opts.lineSearchBeta = o["lineSearchBeta"];
    if (o.count("lineSearchGamma")) opts.lineSearchGamma = o["lineSearchGamma"];
    if (o.count("checkLastActiveSet")) opts.checkLastActiveSet = o["checkLastActiveSet"].is_boolean() ? (int) o["checkLastActiveSet"].get<bool>() : o["checkLastActiveSet"].get<int>();
    opts.regType = (regType_t) reg_type(o);
    if (o.count("regTol")) opts.regTol = o["regTol"];
    if (o.count("regValue")) opts.regValue = o["regValue"];
}

int main(int argc, char **argv)
{
    std::string in_path, init_path, out_path;
    int nrep = 1;
    bool dry = false;
    for (int a = 1; a < argc; a++)
    {
        const std::string s = argv[a];
        if (s == "--out" && a + 1 < argc) out_path = argv[++a];
        else if (s == "--nrep" && a + 1 < argc) nrep = atoi(argv[++a]);
        else if (s == "--dry-run") dry = true;
        else if (in_path.empty()) in_path = s;
        else init_path = s;
    }
    if (in_path.empty()) { fprintf(stderr, "usage: solve_qp_json problem.json [init.json] [--out f] [--nrep N] [--dry-run]\n"); return 2; }
    json jin, jinit, jout;
    { std::ifstream f(in_path); if (!f) { fprintf(stderr, "cannot read %s\n", in_path.c_str()); return 2; } f >> jin; }
    if (!init_path.empty()) { std::ifstream f(init_path); f >> jinit; }
    else if (jin.count("init")) jinit = jin["init"];

    /* ---- dimensions and tree (children counted from the edge list, cpp:236-256) ---- */
    const json &nodes = jin.at("nodes");
    const json &edges = jin.at("edges");
    const int Nn = (int) nodes.size();
    std::vector<int> nx(Nn), nu(Nn), nc(Nn, 0), nk(Nn, 0);
    for (int i = 0; i < Nn; i++) { nx[i] = (int) count_of(nodes[i].at("q")); nu[i] = (int) count_of(nodes[i].at("r")); }
    for (const auto &e : edges) nk.at(e.at("from").get<int>())++;
    if ((int) edges.size() != Nn - 1) throw std::invalid_argument("a tree with N nodes has N-1 edges");

    tree_qp_in qp;
    std::vector<char> qp_mem((size_t) tree_qp_in_calculate_size(Nn, nx.data(), nu.data(), nc.data(), nk.data()));
    tree_qp_in_create(Nn, nx.data(), nu.data(), nc.data(), nk.data(), &qp, qp_mem.data());
    for (const auto &e : edges)
    {
        const int to = e.at("to"), from = e.at("from");
        const dvec A = colmajor(e.at("A"), nx[to], nx[from]), B = colmajor(e.at("B"), nx[to], nu[from]), b = flat(e.at("b"), nx[to]);
        tree_qp_in_set_edge_dynamics_colmajor(A.data(), B.data(), b.data(), &qp, to - 1);
    }
    for (int i = 0; i < Nn; i++)
    {
        const json &n = nodes[i];
        const dvec Q = colmajor(n.at("Q"), nx[i], nx[i]), R = colmajor(n.at("R"), nu[i], nu[i]), S = colmajor(n.at("S"), nu[i], nx[i]);
        const dvec q = flat(n.at("q"), nx[i]), r = flat(n.at("r"), nu[i]);
        tree_qp_in_set_node_objective_colmajor(Q.data(), R.data(), S.data(), q.data(), r.data(), &qp, i);
        const dvec lx = bound(n, "lx", nx[i], -TREEQP_INF), ux = bound(n, "ux", nx[i], TREEQP_INF);
        const dvec lu = bound(n, "lu", nu[i], -TREEQP_INF), uu = bound(n, "uu", nu[i], TREEQP_INF);
        tree_qp_in_set_node_bounds(lx.data(), ux.data(), lu.data(), uu.data(), &qp, i);
    }
    dvec x0(nx[0]);
    if (nx[0] > 0 && jinit.count("x0"))
    {   /* the initial state overrides the root's bounds (cpp:304-312) */
        x0 = flat(jinit.at("x0"), nx[0]);
        tree_qp_in_set_node_bounds(x0.data(), x0.data(), NULL, NULL, &qp, 0);
    }
    bool pinned = nx[0] > 0;
    for (int i = 0; i < nx[0]; i++)
    {
        x0[i] = qp.xmin[0].pa[i];
        pinned = pinned && qp.xmin[0].pa[i] == qp.xmax[0].pa[i];
    }

    tree_qp_out out;
    std::vector<char> out_mem((size_t) tree_qp_out_calculate_size(Nn, nx.data(), nu.data(), nc.data()));
    tree_qp_out_create(Nn, nx.data(), nu.data(), nc.data(), &out, out_mem.data());

    const json opt = jin.count("options") ? jin["options"] : json::object();
    const std::string solver = opt.count("solver") ? opt["solver"].get<std::string>() : "tdunes";
    /* the harness always eliminates a pinned x0 (cpp:350-353); an unpinned root stays a variable */
    if (pinned) tree_qp_in_eliminate_x0(&qp);

    int dim_lam = 0;
    for (int i = 1; i < Nn; i++) dim_lam += nx[i];
    if (dry)
    {
        printf("{\"nodes\": %d, \"solver\": \"%s\", \"nx0_eliminated\": %s, \"dim_lam_tree\": %d}\n", Nn, solver.c_str(), pinned ? "true" : "false", dim_lam);
        return 0;
    }

    int status = -1, iters = -1;
    double best = 1e300;
    std::vector<char> opts_mem, work_mem;
    json ls_iters = json::array();
    if (solver == "tdunes")
    {
        treeqp_tdunes_opts_t o;
        opts_mem.resize((size_t) treeqp_tdunes_opts_calculate_size(Nn));
        treeqp_tdunes_opts_create(Nn, &o, opts_mem.data());
        treeqp_tdunes_opts_set_default(Nn, &o);
        const bool clipping = opt.count("clipping") && (opt["clipping"].is_boolean() ? opt["clipping"].get<bool>() : opt["clipping"].get<int>() != 0);
        for (int i = 0; i < Nn; i++) o.qp_solver[i] = clipping ? TREEQP_CLIPPING_SOLVER : TREEQP_QPOASES_SOLVER; /* cpp:384-394 */
        common_options(opt, o);
        dvec lam0((size_t) dim_lam, 0.0);
        if (jinit.count("lam0_tree")) lam0 = flat(jinit.at("lam0_tree"), dim_lam);
        treeqp_tdunes_workspace work;
        work_mem.resize((size_t) treeqp_tdunes_calculate_size(&qp, &o));
        treeqp_tdunes_create(&qp, &o, &work, work_mem.data());
        for (int rep = 0; rep < nrep; rep++)
        {
            treeqp_tdunes_set_dual_initialization(lam0.data(), &work);
            const auto t0 = std::chrono::steady_clock::now();
            const int st = treeqp_tdunes_solve(&qp, &out, &o, &work);
            const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (rep > 0 && (st != status || out.info.iter != iters)) { fprintf(stderr, "solve %d differs from solve 0\n", rep); return 3; }
            status = st; iters = out.info.iter;
            if (t < best) best = t;
        }
        treeqp_b200_stats stats;
        treeqp_tdunes_get_stats(&work, &stats);
        for (int i = 0; i < iters && i < 256; i++) ls_iters.push_back(stats.ls_iters[i]);
        jout["init"]["lam0_tree"] = lam0;
        treeqp_tdunes_destroy(&work);
    }
    else if (solver == "sdunes")
    {
        if (qp.nx[0] != 0) throw std::invalid_argument("sdunes needs a pinned initial state (it is eliminated)");
        treeqp_sdunes_opts_t o;
        opts_mem.resize((size_t) treeqp_sdunes_opts_calculate_size(Nn));
        treeqp_sdunes_opts_create(Nn, &o, opts_mem.data());
        treeqp_sdunes_opts_set_default(Nn, &o);
        common_options(opt, o);
        treeqp_sdunes_workspace work;
        work_mem.resize((size_t) treeqp_sdunes_calculate_size(&qp, &o));
        treeqp_sdunes_create(&qp, &o, &work, work_mem.data());
        const int nl = treeqp_sdunes_calculate_dual_dimension(work.Nr, work.md, nu[0]);
        const int nm = work.Ns * work.Nh * nx[1];
        dvec lam0((size_t) nl, 0.0), mu0((size_t) nm, 0.0);
        if (jinit.count("lam0_scen")) lam0 = flat(jinit.at("lam0_scen"), nl);
        if (jinit.count("mu0_scen")) mu0 = flat(jinit.at("mu0_scen"), nm);
        for (int rep = 0; rep < nrep; rep++)
        {
            treeqp_sdunes_set_dual_initialization(lam0.data(), mu0.data(), &work);
            const auto t0 = std::chrono::steady_clock::now();
            const int st = treeqp_sdunes_solve(&qp, &out, &o, &work);
            const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (rep > 0 && (st != status || out.info.iter != iters)) { fprintf(stderr, "solve %d differs from solve 0\n", rep); return 3; }
            status = st; iters = out.info.iter;
            if (t < best) best = t;
        }
        jout["init"]["lam0_scen"] = lam0;
        jout["init"]["mu0_scen"] = mu0;
        treeqp_sdunes_destroy(&work);
    }
    else
    {
        fprintf(stderr, "solver \"%s\" is not part of this library (tdunes, sdunes)\n", solver.c_str());
        return 2;
    }

    /* ---- solution in the harness' format (qpSolutionToJson, cpp:112-160) ---- */
    const double kkt = tree_qp_out_max_KKT_res(&qp, &out);
    json sol;
    for (int i = 0; i < Nn; i++)
    {
        json jn;
        const int nxi = (i == 0 && pinned) ? 0 : nx[i];
        dvec x(nxi), mux(nxi), u(nu[i]), muu(nu[i]);
        if (nxi) { tree_qp_out_get_node_x(x.data(), &out, i); tree_qp_out_get_node_mu_x(mux.data(), &out, i); }
        if (nu[i]) { tree_qp_out_get_node_u(u.data(), &out, i); tree_qp_out_get_node_mu_u(muu.data(), &out, i); }
        if (i == 0 && pinned) { x = x0; mux.assign(nx[0], 0.0); }
        jn["x"] = x; jn["mu_x"] = mux; jn["u"] = u; jn["mu_u"] = muu;
        sol["nodes"].push_back(jn);
    }
    for (int e = 0; e < Nn - 1; e++)
    {
        dvec lam(nx[e + 1]);
        tree_qp_out_get_edge_lam(lam.data(), &out, e);
        json je; je["lam"] = lam;
        sol["edges"].push_back(je);
    }
    jout["solution"] = sol;
    jout["info"] = {{"solver", solver}, {"cpu_time", best}, {"status", status}, {"num_iter", iters}, {"kkt_tol", kkt},
                    {"ls_iters", ls_iters}, {"library", treeqp_b200_version()}};
    if (out_path.empty()) std::cout << jout.dump() << std::endl;
    else { std::ofstream f(out_path); f << jout.dump(1) << std::endl; }
    return status == TREEQP_OPTIMAL_SOLUTION_FOUND ? 0 : 1;
}
