//! Runs the programs of oracle/workloads.hpp (SLEEP, PINGPONG = BASELINE config 1, CHAOS, RAFT = BASELINE configs 2
//! and 3) and oracle/workloads_election.hpp (madsim-etcd-client's election API) on the
//! REAL madsim executor and prints, per seed, two observables that pin the whole execution:
//!   elapsed_ns  virtual time at the main future's last poll (every latency draw and every
//!               50-100 ns poll advance before it is folded into this number)
//!   next_u64    the next output of the global RNG at that point (= how many draws were made)
//! `python expected.py --check probe.jsonl` compares them with what the oracle (and therefore the
//! device engine, which is bit-exact against the oracle) produces for the same seeds.
use madsim::net::rpc::Request;
use madsim::net::{Endpoint, NetSim};
use madsim::rand::{random, thread_rng, Rng};
use madsim::runtime::{Handle, Runtime};
use madsim::task::{spawn, NodeId};
use madsim::time::{advance, sleep, Instant};
use madsim_etcd_client::{
    Client, Compare, CompareOp, GetOptions, LeaseGrantOptions, ProclaimOptions, PutOptions, ResignOptions, SimServer, Txn, TxnOp,
};

fn parse_value(v: &[u8]) -> u32 {
    std::str::from_utf8(v).unwrap().parse().unwrap()
}
use serde::{Deserialize, Serialize};
use std::net::SocketAddr;
use std::sync::atomic::{AtomicU32, Ordering};
use std::sync::{Arc, Mutex};
use std::time::Duration;

const SEEDS: [u64; 6] = [0, 1, 2, 42, 2333, 4095];

fn probe(t0: Instant) -> (u64, u64) {
    (t0.elapsed().as_nanos() as u64, random::<u64>())
}

/// oracle/workloads.hpp wl_sleep_main
async fn wl_sleep() -> (u64, u64) {
    let t0 = Instant::now();
    sleep(Duration::from_secs(1)).await;
    probe(t0)
}

/// oracle/workloads.hpp wl_pingpong_main (endpoint.rs:550-584 repeated `rounds` times)
async fn wl_pingpong(rounds: u32) -> (u64, u64) {
    let t0 = Instant::now();
    let handle = Handle::current();
    let addr1: SocketAddr = "10.0.0.1:1".parse().unwrap();
    let node1 = handle.create_node().ip(addr1.ip()).build();
    let node2 = handle.create_node().ip("10.0.0.2".parse().unwrap()).build();
    node1.spawn(async move {
        let ep = Endpoint::bind(addr1).await.unwrap();
        assert_eq!(ep.local_addr().unwrap(), addr1);
        loop {
            let mut buf = vec![0u8; 16];
            let (len, from) = ep.recv_from(1, &mut buf).await.unwrap();
            assert_eq!(&buf[..len], b"ping");
            ep.send_to(from, 1, b"pong").await.unwrap();
        }
    });
    let f = node2.spawn(async move {
        sleep(Duration::from_secs(1)).await;
        let ep = Endpoint::connect(addr1).await.unwrap();
        assert_eq!(ep.peer_addr().unwrap(), addr1);
        let mut done = 0u32;
        for _ in 0..rounds {
            ep.send(1, b"ping").await.unwrap();
            let mut buf = vec![0u8; 16];
            let len = ep.recv(1, &mut buf).await.unwrap();
            assert_eq!(&buf[..len], b"pong");
            done += 1;
            sleep(Duration::from_millis(10)).await;
        }
        done
    });
    assert_eq!(f.await.unwrap(), rounds);
    probe(t0)
}

/// oracle/workloads.hpp wl_chaos_main (task/mod.rs:858-1016 kill / restart / pause_resume in one schedule)
async fn wl_chaos(rounds: u64, period: Duration) -> (u64, u64) {
    let t0 = Instant::now();
    let handle = Handle::current();
    let flag = Arc::new(AtomicU32::new(0));
    let starts = Arc::new(AtomicU32::new(0));
    let (flag1, starts1) = (flag.clone(), starts.clone());
    let node = handle
        .create_node()
        .init(move || {
            let (flag, starts) = (flag1.clone(), starts1.clone());
            async move {
                starts.fetch_add(1, Ordering::Relaxed);
                loop {
                    sleep(period).await;
                    flag.fetch_add(2, Ordering::Relaxed);
                }
            }
        })
        .build();
    let id = node.id();
    for _ in 0..rounds {
        let m: u64 = thread_rng().gen_range(1..4u64);
        sleep(period * (m as u32) + period / 2).await;
        let k: u64 = thread_rng().gen_range(0..3u64);
        if k == 0 {
            handle.kill(id);
            assert!(handle.is_exit(id));
            sleep(period).await;
            handle.restart(id);
            assert!(!handle.is_exit(id));
        } else if k == 1 {
            handle.pause(id);
            let before = flag.load(Ordering::Relaxed);
            sleep(period * 3).await;
            assert_eq!(flag.load(Ordering::Relaxed), before);
            handle.resume(id);
        } else {
            handle.restart(id);
        }
    }
    sleep(period).await;
    probe(t0)
}

// ---------------------------------------------------------------------------------------------
// oracle/workloads.hpp wl_raft_main: 5-node leader election on Endpoint::{call_timeout, add_rpc_handler}
// (BASELINE config 2; with `faults` = 7 the supervisor's partition / clock-jump schedule of config 3).
// MadRaft is not part of the madsim tree: this is the program this repository's oracle and device engine run.
#[derive(Serialize, Deserialize)]
struct RequestVote {
    term: u32,
    candidate: u32,
}
impl Request for RequestVote {
    type Response = (u32, bool);
    const ID: u64 = 0x7261_6674_0001; // any two distinct ids: only tag equality is observable
}
#[derive(Serialize, Deserialize)]
struct AppendEntries {
    term: u32,
    leader: u32,
}
impl Request for AppendEntries {
    type Response = (u32, bool);
    const ID: u64 = 0x7261_6674_0002;
}
#[derive(Default)]
struct Raft {
    term: u32,
    voted_for: u32, // node index, 0 = none
    role: u32,      // 0 follower, 1 candidate, 2 leader
    votes: u32,
    leader: u32,
    last_hb: Option<Instant>,
}
#[derive(Clone)]
struct RaftParams {
    n: u32,
    dur_ms: u64,
    fault_mask: u32,
    rpc_timeout: Duration,
    heartbeat: Duration,
    election_ms: (u64, u64),
}
type Cluster = Arc<Mutex<Vec<Option<Arc<Mutex<Raft>>>>>>; // by node index 1..=n
fn raft_addr(i: u32) -> SocketAddr {
    format!("10.0.0.{i}:1").parse().unwrap()
}
fn step_down(r: &mut Raft, term: u32) {
    r.term = term;
    r.role = 0;
    r.voted_for = 0;
}
async fn raft_append_task(ep: Endpoint, r: Arc<Mutex<Raft>>, p: RaftParams, me: u32, peer: u32, term0: u32) {
    let rsp = ep.call_timeout(raft_addr(peer), AppendEntries { term: term0, leader: me }, p.rpc_timeout).await;
    if let Ok((term, _)) = rsp {
        let mut r = r.lock().unwrap();
        if term > r.term {
            step_down(&mut r, term);
        }
    }
}
async fn raft_leader_loop(ep: Endpoint, r: Arc<Mutex<Raft>>, p: RaftParams, me: u32, term0: u32) {
    loop {
        {
            let r = r.lock().unwrap();
            if !(r.role == 2 && r.term == term0) {
                return;
            }
        }
        for peer in 1..=p.n {
            if peer != me {
                spawn(raft_append_task(ep.clone(), r.clone(), p.clone(), me, peer, term0));
            }
        }
        sleep(p.heartbeat).await;
    }
}
async fn raft_vote_task(ep: Endpoint, r: Arc<Mutex<Raft>>, p: RaftParams, me: u32, peer: u32, term0: u32) {
    let rsp = ep.call_timeout(raft_addr(peer), RequestVote { term: term0, candidate: me }, p.rpc_timeout).await;
    if let Ok((term, granted)) = rsp {
        let mut won = false;
        {
            let mut r = r.lock().unwrap();
            if term > r.term {
                step_down(&mut r, term);
            } else if r.role == 1 && r.term == term0 && granted {
                r.votes += 1;
                if r.votes == p.n / 2 + 1 {
                    r.role = 2;
                    r.leader = me;
                    won = true;
                }
            }
        }
        if won {
            spawn(raft_leader_loop(ep.clone(), r.clone(), p.clone(), me, term0));
        }
    }
}
async fn raft_node_main(cluster: Cluster, p: RaftParams, me: u32) {
    let ep = Endpoint::bind(raft_addr(me)).await.unwrap();
    let r = Arc::new(Mutex::new(Raft { last_hb: Some(Instant::now()), ..Default::default() }));
    cluster.lock().unwrap()[me as usize] = Some(r.clone());
    let rv = r.clone();
    ep.add_rpc_handler(move |req: RequestVote| {
        let r = rv.clone();
        async move {
            let mut r = r.lock().unwrap();
            if req.term > r.term {
                step_down(&mut r, req.term);
            }
            let granted = req.term == r.term && (r.voted_for == 0 || r.voted_for == req.candidate);
            if granted {
                r.voted_for = req.candidate;
                r.last_hb = Some(Instant::now());
            }
            (r.term, granted)
        }
    });
    let ra = r.clone();
    ep.add_rpc_handler(move |req: AppendEntries| {
        let r = ra.clone();
        async move {
            let mut r = r.lock().unwrap();
            let ok = req.term >= r.term;
            if ok {
                if req.term > r.term {
                    r.term = req.term;
                    r.voted_for = 0;
                }
                r.role = 0;
                r.leader = req.leader;
                r.last_hb = Some(Instant::now());
            }
            (r.term, ok)
        }
    });
    loop {
        let to_ms: u64 = thread_rng().gen_range(p.election_ms.0..p.election_ms.1);
        sleep(Duration::from_millis(to_ms)).await;
        let mut term0 = None;
        {
            let mut r = r.lock().unwrap();
            if r.role != 2 && r.last_hb.unwrap().elapsed() >= Duration::from_millis(to_ms) {
                r.term += 1;
                r.role = 1;
                r.voted_for = me;
                r.votes = 1;
                r.last_hb = Some(Instant::now());
                term0 = Some(r.term);
            }
        }
        if let Some(term0) = term0 {
            for peer in 1..=p.n {
                if peer != me {
                    spawn(raft_vote_task(ep.clone(), r.clone(), p.clone(), me, peer, term0));
                }
            }
        }
    }
}
/// oracle/workloads.hpp raft_fault_task: every choice is drawn from the seed's own RNG
async fn raft_fault_task(p: RaftParams, ids: Vec<NodeId>) {
    let net = NetSim::current();
    let kinds: Vec<u32> = (0..3).filter(|k| p.fault_mask & (1 << k) != 0).collect();
    let id = |i: u64| ids[(i - 1) as usize];
    loop {
        let nap: u64 = thread_rng().gen_range(500..2000u64);
        sleep(Duration::from_millis(nap)).await;
        let k = kinds[thread_rng().gen_range(0..kinds.len() as u64) as usize];
        if k == 0 {
            let i: u64 = thread_rng().gen_range(1..p.n as u64 + 1);
            net.clog_node(id(i));
            let d: u64 = thread_rng().gen_range(200..1000u64);
            sleep(Duration::from_millis(d)).await;
            net.unclog_node(id(i));
        } else if k == 1 {
            let i: u64 = thread_rng().gen_range(1..p.n as u64 + 1);
            let mut j: u64 = thread_rng().gen_range(1..p.n as u64);
            if j >= i {
                j += 1;
            }
            net.clog_link(id(i), id(j));
            net.clog_link(id(j), id(i));
            let d: u64 = thread_rng().gen_range(200..1000u64);
            sleep(Duration::from_millis(d)).await;
            net.unclog_link(id(i), id(j));
            net.unclog_link(id(j), id(i));
        } else {
            let d: u64 = thread_rng().gen_range(0..500u64);
            advance(Duration::from_millis(d));
        }
    }
}
async fn wl_raft(fault_mask: u32) -> (u64, u64) {
    let t0 = Instant::now();
    let handle = Handle::current();
    let p = RaftParams {
        n: 5,
        dur_ms: 10_000,
        fault_mask,
        rpc_timeout: Duration::from_millis(40),
        heartbeat: Duration::from_millis(50),
        election_ms: (150, 300),
    };
    let cluster: Cluster = Arc::new(Mutex::new(vec![None; 6]));
    let mut ids = Vec::new();
    for i in 1..=p.n {
        let (c, pp) = (cluster.clone(), p.clone());
        let node = handle
            .create_node()
            .ip(raft_addr(i).ip())
            .init(move || raft_node_main(c.clone(), pp.clone(), i))
            .build();
        ids.push(node.id());
    }
    if fault_mask != 0 {
        spawn(raft_fault_task(p.clone(), ids));
    }
    for tick in 1..=p.dur_ms / 100 {
        sleep(Duration::from_millis(100)).await;
        let rafts = cluster.lock().unwrap();
        let mut any_leader = false;
        for a in 1..=p.n as usize {
            let Some(ra) = &rafts[a] else { continue };
            let ra = ra.lock().unwrap();
            if ra.role != 2 {
                continue;
            }
            any_leader = true;
            for b in a + 1..=p.n as usize {
                if let Some(rb) = &rafts[b] {
                    let rb = rb.lock().unwrap();
                    assert!(!(rb.role == 2 && rb.term == ra.term), "two leaders in one term");
                }
            }
        }
        assert!(fault_mask != 0 || tick != 20 || any_leader, "no leader elected by 2 s");
    }
    probe(t0)
}

// ---------------------------------------------------------------------------------------------
// oracle/workloads_election.hpp wl_election_main: madsim-etcd-client's election API driven the way the crate's own
// election() test drives it (a first leader, a second candidate parked in campaign(), an observer).
async fn first_leader(hold: Duration) {
    let client = Client::connect(["10.0.0.1:2379"], None).await.unwrap();
    let mut leases = client.lease_client();
    let mut el = client.election_client();
    sleep(Duration::from_secs(5)).await;
    let lease = leases.grant(60, None).await.unwrap();
    let won = el.campaign("leader", "1", lease.id()).await.unwrap();
    let key = won.leader().unwrap().clone();
    assert_eq!(el.leader("leader").await.unwrap().kv().unwrap().value(), b"1");
    el.campaign("leader", "1", lease.id()).await.unwrap();
    el.campaign("leader", "1.1", lease.id()).await.unwrap();
    let opt = ProclaimOptions::new().with_leader(key);
    el.proclaim("1.2", Some(opt.clone())).await.unwrap();
    assert_eq!(el.leader("leader").await.unwrap().kv().unwrap().value(), b"1.2");
    sleep(hold).await;
    leases.revoke(lease.id()).await.unwrap();
    el.proclaim("1.3", Some(opt)).await.unwrap_err();
    el.campaign("invalid_lease", "1", 1).await.unwrap_err();
}
async fn second_leader() {
    let client = Client::connect(["10.0.0.1:2379"], None).await.unwrap();
    let mut leases = client.lease_client();
    let mut el = client.election_client();
    sleep(Duration::from_secs(10)).await;
    let lease = leases.grant(60, None).await.unwrap();
    let won = el.campaign("leader", "2", lease.id()).await.unwrap();
    let key = won.leader().unwrap().clone();
    el.resign(Some(ResignOptions::new().with_leader(key))).await.unwrap();
}
async fn observer() {
    let client = Client::connect(["10.0.0.1:2379"], None).await.unwrap();
    let mut kv = client.kv_client();
    let mut el = client.election_client();
    let mut stream = el.observe("leader").await.unwrap();
    for want in [&b"1"[..], b"1.1", b"1.2"] {
        assert_eq!(stream.message().await.unwrap().unwrap().kv().unwrap().value(), want);
    }
    sleep(Duration::from_secs(15)).await;
    let got = kv.get("leader", Some(GetOptions::new().with_prefix())).await.unwrap();
    assert_eq!(got.kvs().len(), 2);
    assert_eq!(stream.message().await.unwrap().unwrap().kv().unwrap().value(), b"2");
}
async fn wl_election(hold_s: u64) -> (u64, u64) {
    let t0 = Instant::now();
    let handle = Handle::current();
    let node = |ip: &str| handle.create_node().ip(ip.parse().unwrap()).build();
    let (server, c1, c2, c3) = (node("10.0.0.1"), node("10.0.0.2"), node("10.0.0.3"), node("10.0.0.4"));
    server.spawn(async move {
        SimServer::builder().serve("10.0.0.1:2379".parse().unwrap()).await.unwrap();
    });
    sleep(Duration::from_secs(1)).await;
    let t1 = c1.spawn(first_leader(Duration::from_secs(hold_s)));
    let t2 = c2.spawn(second_leader());
    let t3 = c3.spawn(observer());
    t1.await.unwrap();
    t2.await.unwrap();
    t3.await.unwrap();
    probe(t0)
}

// ---------------------------------------------------------------------------------------------
// oracle/workloads_etcd.hpp wl_etcd_main with its default parameters (BASELINE config 4 at loss 0: 2 clients x 32
// operations, 50 ms think time, lease 7 with ttl 1 granted by client 0, no injected timeouts): the crate's own SimServer
// and KvClient / LeaseClient, one connection per request.  Which operation a client issues is drawn from the global RNG
// exactly as the oracle's etcd_client does (gen_range(0..100), gen_range(0..5)); what the store answers does not move
// the two numbers this probe prints (every answer is one message), the number and timing of requests does.
const ETCD_NKEY: u64 = 5;
const ETCD_LEASE_KEY: u64 = 4;
const ETCD_LEASE_ID: i64 = 7;
async fn etcd_client(c: u32, ops: u64, think: Duration) {
    let client = Client::connect(["10.0.0.1:2379"], None).await.unwrap();
    let mut kv = client.kv_client();
    let mut leases = client.lease_client();
    let mut last_seen = [0u32; ETCD_NKEY as usize];
    let mut need_grant = c == 0;
    for i in 0..ops {
        let r = thread_rng().gen_range(0..100u64);
        let key = thread_rng().gen_range(0..ETCD_NKEY);
        let name = format!("k{key}");
        let value = (c + 1) << 12 | (i as u32 + 1);
        if need_grant {
            let opt = LeaseGrantOptions::new().with_id(ETCD_LEASE_ID);
            if leases.grant(1, Some(opt)).await.is_ok() {
                need_grant = false;
            }
        } else if key == ETCD_LEASE_KEY {
            if r < 50 {
                let opt = PutOptions::new().with_lease(ETCD_LEASE_ID);
                match kv.put(name, value.to_string(), Some(opt)).await {
                    Ok(_) => last_seen[key as usize] = value,
                    Err(_) => need_grant = c == 0, // lease not found: client 0 grants it again
                }
            } else if let Ok(got) = kv.get(name, None).await {
                last_seen[key as usize] = got.kvs().first().map_or(0, |kv| parse_value(kv.value()));
            }
        } else if r < 40 {
            if kv.put(name, value.to_string(), None).await.is_ok() {
                last_seen[key as usize] = value;
            }
        } else if r < 75 {
            if let Ok(got) = kv.get(name, None).await {
                last_seen[key as usize] = got.kvs().first().map_or(0, |kv| parse_value(kv.value()));
            }
        } else if r < 85 {
            if kv.delete(name, None).await.is_ok() {
                last_seen[key as usize] = 0;
            }
        } else {
            // compare-and-swap on the value this client saw last ("0" for a key it believes absent: the crate's Equal never
            // matches a missing key, the oracle's model lets it; either way the answer is one message)
            let expect = last_seen[key as usize];
            let txn = Txn::new()
                .when([Compare::value(name.clone(), CompareOp::Equal, expect.to_string())])
                .and_then([TxnOp::put(name.clone(), value.to_string(), None)])
                .or_else([TxnOp::get(name, None)]);
            if let Ok(rsp) = kv.txn(txn).await {
                if rsp.succeeded() {
                    last_seen[key as usize] = value;
                }
            }
        }
        sleep(think).await;
    }
}
async fn wl_etcd(n_clients: u32, ops: u64, think: Duration) -> (u64, u64) {
    let t0 = Instant::now();
    let handle = Handle::current();
    let node = |ip: String| handle.create_node().ip(ip.parse().unwrap()).build();
    let server = node("10.0.0.1".into());
    let clients: Vec<_> = (0..n_clients).map(|c| node(format!("10.0.0.{}", 2 + c))).collect();
    server.spawn(async move {
        SimServer::builder().serve("10.0.0.1:2379".parse().unwrap()).await.unwrap();
    });
    sleep(Duration::from_secs(1)).await;
    let tasks: Vec<_> = clients.iter().enumerate().map(|(c, n)| n.spawn(etcd_client(c as u32, ops, think))).collect();
    for t in tasks {
        t.await.unwrap();
    }
    probe(t0)
}

fn main() {
    for seed in SEEDS {
        let run = |name: &str, out: (u64, u64)| {
            println!(
                "{{\"workload\":\"{name}\",\"seed\":{seed},\"elapsed_ns\":{},\"next_u64\":{}}}",
                out.0, out.1
            );
        };
        let rt = Runtime::with_seed_and_config(seed, madsim::Config::default());
        run("sleep", rt.block_on(wl_sleep()));
        let rt = Runtime::with_seed_and_config(seed, madsim::Config::default());
        run("pingpong", rt.block_on(wl_pingpong(1000)));
        let rt = Runtime::with_seed_and_config(seed, madsim::Config::default());
        run("chaos", rt.block_on(wl_chaos(8, Duration::from_millis(2000))));
        let rt = Runtime::with_seed_and_config(seed, madsim::Config::default());
        run("raft", rt.block_on(wl_raft(0))); // BASELINE config 2
        let rt = Runtime::with_seed_and_config(seed, madsim::Config::default());
        run("raft_faults", rt.block_on(wl_raft(7))); // BASELINE config 3
        let rt = Runtime::with_seed_and_config(seed, madsim::Config::default());
        run("election", rt.block_on(wl_election(30)));
        let rt = Runtime::with_seed_and_config(seed, madsim::Config::default());
        run("etcd", rt.block_on(wl_etcd(2, 32, Duration::from_millis(50)))); // BASELINE config 4, loss 0
    }
}
