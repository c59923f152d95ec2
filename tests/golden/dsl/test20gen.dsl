--
-- Query in SQL:
-- SELECT col1 FROM tbl4_ctrl WHERE col3 >= 139 and col3 < 140;
-- SELECT col1 FROM tbl4_ctrl WHERE col3 >= 75 and col3 < 77;
--
s1=select(db1.tbl4_ctrl.col3,139,140)
f1=fetch(db1.tbl4_ctrl.col1,s1)
print(f1)
s2=select(db1.tbl4_ctrl.col3,75,77)
f2=fetch(db1.tbl4_ctrl.col1,s2)
print(f2)
